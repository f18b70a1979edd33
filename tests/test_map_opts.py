"""lra_b200_map_opts: the Python mirror has the library's layout, and the four align presets carry the values of lra.cpp:268-431 on top of Options.h:123-240."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from lra_b200 import capi  # noqa: E402


def test_mirror_has_the_library_layout():
    import emu_mp
    L = emu_mp.lib(1)
    assert L.emu_sizeof_mpopts() == C.sizeof(capi.MapOpts)


def test_align_presets():
    want = {
        "ont": dict(globalK=17, globalW=10, globalMaxFreq=150, readType=0, NumAln=2, bypassClustering=1, HighlyAccurate=0, refineBand=7, cleanMaxDiag=200, minDiagCluster=3),
        "clr": dict(globalK=15, globalMaxFreq=250, readType=1, NumAln=2, bypassClustering=1, HighlyAccurate=0, refineBand=20, SecondCleanMaxDiag=120),
        # lra.cpp:306-338
        "ccs": dict(globalK=25, globalW=20, globalMaxFreq=150, readType=2, NumAln=2, bypassClustering=0, HighlyAccurate=1, refineBand=7, merge_dist=100, RoughClustermaxGap=500,
                    maxGap=400, maxDiag=500, cleanMaxDiag=150, SecondCleanMaxDiag=100, SecondCleanMinDiagCluster=30, minDiagCluster=10, minClusterSize=10, refineSpaceDist=30000,
                    hardClip=1, gapCeiling1=2000, gapCeiling2=3000, minUniqueStretchNum=1, minUniqueStretchDist=50),
        # lra.cpp:268-305
        "contig": dict(globalK=19, globalW=10, globalMaxFreq=30, readType=3, NumAln=2, bypassClustering=0, HighlyAccurate=1, refineBand=50, maxDiag=100, maxGap=500,
                       RoughClustermaxGap=500, minDiagCluster=30, minClusterSize=10, refineSpaceDist=50000, gapCeiling1=3000, gapCeiling2=5000, merge_dist=100),
    }
    for preset, fields in want.items():
        o = capi.map_opts_preset(preset)
        for k, v in fields.items():
            assert getattr(o, k) == v, (preset, k, getattr(o, k), v)
    o = capi.map_opts_preset("ccs")
    assert abs(o.initial_anchorbonus - 10.0) < 1e-6 and abs(o.gapextend - 15.0) < 1e-6 and abs(o.gapopen - 4.0) < 1e-6 and abs(o.anchorstoosparse - 0.005) < 1e-9
    import pytest
    with pytest.raises(capi.LraB200Error):
        capi.map_opts_preset("nanopore")
