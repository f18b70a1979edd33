"""lra_b200 -- B200-native drop-in for the MapRead hot path of ChaissonLab/LRA.

The product is the C-ABI shared library `liblra_b200.so` (include/lra_b200.h, sources in lra_b200/csrc/).  This package
is the thin Python host-side mirror used by the tests and bench.py: same names and argument meaning as the reference's
functions for this path.  There is no CPU implementation here: importing works anywhere, but every compute call needs
the CUDA library and a GPU and raises otherwise.
"""
from .capi import (  # noqa: F401
    Context, LraB200Error, SeqArena, LocalIndexImage, CreateLookUpTable, write_gli, read_gli, split_chain_view, AffineOneGapAlign, AffineOneGapAlignBatch, IndelRefineAlignment, library_path, load_library, init_pwl, map_opts_preset, read_mms, format_sam, Mapper, MapOpts, RECORD, write_mms, build_index_files, INDEX_PRESET,
)
