"""ctypes binding of include/lra_b200.h plus the host-side mirror of the reference interface for the hot path."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LraB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lra_b200 error %d: %s" % (code, msg))
        self.code = code


OK, EINVAL, ECUDA, EOVERFLOW, EINTERNAL = 0, 1, 2, 3, 4


class _AogJobs(C.Structure):
    _fields_ = [("q_off", C.c_void_p), ("t_off", C.c_void_p), ("q_len", C.c_void_p), ("t_len", C.c_void_p),
                ("k", C.c_void_p), ("n_jobs", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32),
                ("indel", C.c_int32)]


class _AogResult(C.Structure):
    _fields_ = [("score", C.c_void_p), ("n_blocks", C.c_void_p), ("block_off", C.c_void_p), ("blocks", C.c_void_p),
                ("block_cap", C.c_uint64), ("n_blocks_total", C.c_uint64), ("cells", C.c_uint64)]


class _IrGroups(C.Structure):
    _fields_ = [("q_base", C.c_void_p), ("t_base", C.c_void_p), ("q_start", C.c_void_p), ("t_start", C.c_void_p),
                ("t_len", C.c_void_p), ("q_seq_len", C.c_void_p), ("t_seq_len", C.c_void_p), ("band_off", C.c_void_p),
                ("band", C.c_void_p), ("band_len", C.c_uint64), ("n_groups", C.c_int32), ("match", C.c_int32),
                ("mismatch", C.c_int32), ("indel", C.c_int32)]


class _IrResult(C.Structure):
    _fields_ = [("n_blocks", C.c_void_p), ("block_off", C.c_void_p), ("blocks", C.c_void_p), ("block_cap", C.c_uint64),
                ("n_blocks_total", C.c_uint64), ("cells", C.c_uint64)]


class _IrSegments(C.Structure):
    _fields_ = [("blocks_in", C.c_void_p), ("blk_off", C.c_void_p), ("blk_cnt", C.c_void_p), ("q_base", C.c_void_p),
                ("t_base", C.c_void_p), ("read_len", C.c_void_p), ("contig_len", C.c_void_p), ("n_blocks_in", C.c_uint64),
                ("n_segments", C.c_int32), ("refine_band", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32),
                ("indel", C.c_int32), ("end_align", C.c_int32)]


class _IrSegResult(C.Structure):
    _fields_ = [("n_blocks", C.c_void_p), ("block_off", C.c_void_p), ("blocks", C.c_void_p), ("block_cap", C.c_uint64),
                ("n_blocks_total", C.c_uint64), ("cells", C.c_uint64), ("n_dp_groups", C.c_uint64), ("n_aog_jobs", C.c_uint64)]


class _SdpProblems(C.Structure):
    _fields_ = [("n_prob", C.c_int32), ("max_aln", C.c_int32), ("mode", C.c_void_p), ("frag_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("len", C.c_void_p),
                ("cl_off_off", C.c_void_p), ("cl_off", C.c_void_p), ("cl_strand", C.c_void_p), ("only_cl", C.c_void_p), ("rate", C.c_void_p), ("irate", C.c_void_p),
                ("read_len", C.c_void_p), ("alnthres", C.c_float), ("num_aln", C.c_int32), ("pwl_stops", C.c_void_p), ("pwl_slope", C.c_void_p), ("pwl_inter", C.c_void_p),
                ("ceil1", C.c_int32), ("ceil2", C.c_int32), ("q_end", C.c_void_p), ("t_end", C.c_void_p), ("frag_strand", C.c_void_p), ("frag_val", C.c_void_p),
                ("frag_n0", C.c_void_p), ("global_k", C.c_int32)]


class _SdpResult(C.Structure):
    _fields_ = [("n_chains", C.c_void_p), ("chain_len", C.c_void_p), ("chain_val", C.c_void_p), ("bounds", C.c_void_p), ("chain", C.c_void_p), ("link", C.c_void_p),
                ("cl_of_frag", C.c_void_p), ("arena_peak", C.c_uint64), ("num_anchors0", C.c_void_p)]


class MapOpts(C.Structure):
    """lra_b200_map_opts (include/lra_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in ("globalK", "globalW", "globalMaxFreq", "localW", "localMaxFreq", "smallK", "smallW", "cleanMaxDiag", "minDiagCluster",
                                         "cleanClustersize", "SecondCleanMinDiagCluster", "SecondCleanMaxDiag", "punish_anchorfreq", "anchorPerlength", "NumAln",
                                         "PrintNumAln", "splitdist", "readType")] + \
               [(n, C.c_float) for n in ("initial_anchorbonus", "second_anchorbonus", "alnthres", "anchorstoosparse")] + \
               [(n, C.c_int32) for n in ("refineSpaceDist", "window", "limitrefine", "RefineBySDP", "localMatch", "localMismatch", "localIndel", "localBand", "refineBand",
                                         "hardClip", "bypassClustering")] + \
               [(n, C.c_float) for n in ("gapopen", "gapextend", "gaproot")] + \
               [(n, C.c_int32) for n in ("gapCeiling1", "gapCeiling2", "localIndexWindow", "localIndexMaxFreq", "HighlyAccurate", "maxDiag", "maxGap", "RoughClustermaxGap",
                                         "minClusterSize", "minUniqueStretchNum", "minUniqueStretchDist", "merge_dist")]


# lra_b200_record
RECORD = np.dtype([(n, "<i4") for n in ("read", "chain", "seg", "n_seg")] + [("flag", "<u4")] +
                  [(n, "<i4") for n in ("chrom", "strand", "mapq", "order", "typeofaln", "supplementary")] +
                  [(n, "<u4") for n in ("tStart", "tEnd", "qStart", "qEnd")] + [(n, "<i4") for n in ("preClip", "sufClip")] +
                  [(n, "<i4") for n in ("nm", "nmm", "nins", "ndel", "tins", "tdel", "nSmallDel", "nMedDel", "nLargeDel", "nSmallIns", "nMedIns", "nLargeIns")] +
                  [("value", "<f4")] + [(n, "<i4") for n in ("NumOfAnchors0", "NumOfAnchors1", "n_blocks", "n_cigar")] + [("cigar_off", "<u8")])


class _MapResult(C.Structure):
    _fields_ = [("status", C.c_void_p), ("n_aln", C.c_void_p), ("aln_nseg", C.c_void_p), ("aln_seg0", C.c_void_p), ("aln_rank", C.c_void_p), ("records", C.c_void_p),
                ("record_cap", C.c_uint64), ("n_records", C.c_uint64), ("cigar", C.c_void_p), ("cigar_cap", C.c_uint64), ("n_cigar", C.c_uint64), ("aligned_bases", C.c_uint64)]


class _SeedReads(C.Structure):
    _fields_ = [("read_off", C.c_void_p), ("read_len", C.c_void_p), ("n_reads", C.c_int32), ("k", C.c_int32), ("w", C.c_int32),
                ("max_freq", C.c_int64)]


class _SeedResult(C.Structure):
    _fields_ = [("match_off", C.c_void_p), ("q_t", C.c_void_p), ("q_pos", C.c_void_p), ("t_t", C.c_void_p), ("t_pos", C.c_void_p),
                ("strand", C.c_void_p), ("match_cap", C.c_uint64), ("n_matches", C.c_uint64), ("n_minimizers", C.c_void_p)]


class _StatsResult(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("value", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("cigar_cap", C.c_uint64),
                ("n_cigar_total", C.c_uint64)]


class _Clusters(C.Structure):
    _fields_ = [("n_clusters", C.c_int32), ("m_q", C.c_void_p), ("m_t", C.c_void_p), ("m_off", C.c_void_p), ("box", C.c_void_p),
                ("strand", C.c_void_p), ("read_id", C.c_void_p), ("hdr_pos", C.c_void_p), ("n_hdr", C.c_int32), ("global_k", C.c_int32),
                ("small_k", C.c_int32), ("window", C.c_int32), ("local_max_freq", C.c_int64)]


class _SplitChains(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("m_q", C.c_void_p), ("m_t", C.c_void_p), ("m_len", C.c_void_p), ("m_strand", C.c_void_p), ("m_off", C.c_void_p),
                ("box", C.c_void_p), ("strand", C.c_void_p), ("chrom", C.c_void_p), ("read_id", C.c_void_p), ("hdr_pos", C.c_void_p), ("n_hdr", C.c_int32),
                ("global_k", C.c_int32), ("small_k", C.c_int32), ("window", C.c_int32), ("local_max_freq", C.c_int64), ("limitrefine", C.c_int32)]


class _ExtendParts(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("g_off", C.c_void_p), ("p_off", C.c_void_p), ("p_strand", C.c_void_p), ("chrom_off", C.c_void_p), ("chrom_len", C.c_void_p),
                ("read_off", C.c_void_p), ("read_len", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("K", C.c_int32), ("skipsorting", C.c_int32), ("trim", C.c_int32)]


class _Extended(C.Structure):
    _fields_ = [("e_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("len", C.c_void_p), ("cap", C.c_uint64), ("n_total", C.c_uint64), ("box", C.c_void_p)]


class _ExtendChains(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("ch_off", C.c_void_p), ("ch", C.c_void_p), ("n_clusters", C.c_int32), ("cl_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p),
                ("cl_box", C.c_void_p), ("cl_strand", C.c_void_p), ("cl_freq", C.c_void_p), ("chrom_off", C.c_void_p), ("chrom_len", C.c_void_p), ("read_off", C.c_void_p),
                ("read_len", C.c_void_p), ("K", C.c_int32), ("skiprepetitive", C.c_int32), ("trim", C.c_int32), ("merge_dist", C.c_int32)]


class _ExtendedChains(C.Structure):
    _fields_ = [("e_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("len", C.c_void_p), ("ovp", C.c_void_p), ("md_head", C.c_void_p), ("cap", C.c_uint64),
                ("n_total", C.c_uint64), ("box", C.c_void_p), ("overlap", C.c_void_p)]


class _AnchorChains(C.Structure):
    _fields_ = [("n_chains", C.c_int32), ("c_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("len", C.c_void_p), ("strand", C.c_void_p), ("cnum", C.c_void_p),
                ("link", C.c_void_p), ("hdr_pos", C.c_void_p), ("n_hdr", C.c_int32), ("splitdist", C.c_int32), ("bypass_clustering", C.c_int32)]


class _SplitChainsOut(C.Structure):
    _fields_ = [("n_sp", C.c_void_p), ("n_link", C.c_void_p), ("sp_off", C.c_void_p), ("ci_off", C.c_void_p), ("sptc", C.c_void_p), ("ci", C.c_void_p), ("sp_lk", C.c_void_p),
                ("sp_box", C.c_void_p), ("sp_chrom", C.c_void_p), ("sp_type", C.c_void_p), ("sp_strand", C.c_void_p), ("sp_link", C.c_void_p)]


def split_chain_view(o, c_off, k):
    """Chain k of a split_chains_batch result as the per-chain dict the oracle returns (sp_off, sptc, sp_lk, ci_off, ci, box, chrom, type, strand, link)."""
    a = int(c_off[k]); ns = int(o["n_sp"][k])
    so = o["sp_off"][a + k:a + k + ns + 1]; co = o["ci_off"][a + k:a + k + ns + 1]
    m = int(so[ns]) if ns else 0; mc = int(co[ns]) if ns else 0
    return dict(sp_off=so, sptc=o["sptc"][a:a + m], sp_lk=o["sp_lk"][a:a + m], ci_off=co, ci=o["ci"][a:a + mc], box=o["sp_box"][a:a + ns], chrom=o["sp_chrom"][a:a + ns],
                type=o["sp_type"][a:a + ns], strand=o["sp_strand"][a:a + ns], link=o["sp_link"][a:a + int(o["n_link"][k])])


class _LinearGaps(C.Structure):
    _fields_ = [("n_gaps", C.c_int32), ("cur_read_end", C.c_void_p), ("next_read_start", C.c_void_p), ("cur_genome_end", C.c_void_p), ("next_genome_start", C.c_void_p),
                ("read_off", C.c_void_p), ("chrom_off", C.c_void_p), ("match", C.c_int32), ("mismatch", C.c_int32), ("indel", C.c_int32), ("local_band", C.c_int32)]


class _Spaces(C.Structure):
    _fields_ = [("n_spaces", C.c_int32), ("qs", C.c_void_p), ("qe", C.c_void_p), ("ts", C.c_void_p), ("te", C.c_void_p), ("lrts", C.c_void_p), ("lrlength", C.c_void_p),
                ("read_off", C.c_void_p), ("read_len", C.c_void_p), ("chrom_off", C.c_void_p), ("flip", C.c_void_p), ("K", C.c_int32), ("match", C.c_int32),
                ("mismatch", C.c_int32), ("indel", C.c_int32), ("W", C.c_int32), ("local_max_freq", C.c_int32), ("diag", C.c_void_p)]


class _SpaceResult(C.Structure):
    _fields_ = [("pair_off", C.c_void_p), ("n_pairs", C.c_void_p), ("identity", C.c_void_p), ("pq", C.c_void_p), ("pt", C.c_void_p), ("pair_cap", C.c_uint64),
                ("n_pairs_total", C.c_uint64)]


class _RoughLists(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("l_off", C.c_void_p), ("lr_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("r_start", C.c_void_p), ("r_end", C.c_void_p),
                ("r_box", C.c_void_p), ("r_strand", C.c_void_p), ("r_freq", C.c_void_p), ("r_chrom", C.c_void_p), ("globalK", C.c_int32), ("max_gap", C.c_int32),
                ("min_cluster_size", C.c_int32), ("max_diag", C.c_int32)]


class _SplitRoughResult(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ["n_split", "n_piece", "s_start", "s_end", "s_coarse", "s_chrom", "s_box", "s_strand", "s_freq", "p_cluster", "p_start", "p_end"]]


class _CleanedLists(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("l_off", C.c_void_p), ("q", C.c_void_p), ("t", C.c_void_p), ("qt", C.c_void_p), ("freq", C.c_void_p), ("strand", C.c_void_p),
                ("hdr_pos", C.c_void_p), ("n_hdr", C.c_int32), ("globalK", C.c_int32), ("max_diag", C.c_int32), ("min_cluster_size", C.c_int32),
                ("min_cluster_length", C.c_int32), ("bypass_clustering", C.c_int32)]


class _DiagClusters(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ["n_cl", "c_start", "c_end", "c_chrom", "c_box", "c_freq"]]


class _Breakpoints(C.Structure):
    _fields_ = [("n_pairs", C.c_int32), ("lf", C.c_void_p), ("ll", C.c_void_p), ("rf", C.c_void_p), ("rl", C.c_void_p), ("lstrand", C.c_void_p),
                ("rstrand", C.c_void_p), ("read_off", C.c_void_p), ("read_len", C.c_void_p), ("lchrom_off", C.c_void_p), ("rchrom_off", C.c_void_p),
                ("lchrom_len", C.c_void_p), ("rchrom_len", C.c_void_p)]


class _BreakpointResult(C.Structure):
    _fields_ = [("mode", C.c_void_p), ("n_out", C.c_void_p), ("bound", C.c_void_p), ("out", C.c_void_p), ("refined", C.c_void_p)]


class _AnchorLists(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("q", C.c_void_p), ("t", C.c_void_p), ("qt", C.c_void_p), ("list_off", C.c_void_p), ("strand", C.c_void_p),
                ("hdr_pos", C.c_void_p), ("n_hdr", C.c_int32)]


class _CleanOpts(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ["cleanMaxDiag", "minDiagCluster", "bypassClustering", "cleanClustersize", "SecondCleanMinDiagCluster", "punish_anchorfreq",
                                         "anchorPerlength", "SecondCleanMaxDiag", "ExtractDiagonalFromClean", "globalK"]]


class _CleanResult(C.Structure):
    _fields_ = [("keep", C.c_void_p), ("freq", C.c_void_p), ("cnt", C.c_void_p), ("cl", C.c_void_p), ("cl_freq", C.c_void_p), ("n_cl", C.c_void_p)]


class _ReadClusters(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("cl_off", C.c_void_p), ("box", C.c_void_p), ("strand", C.c_void_p), ("freq", C.c_void_p), ("m_off", C.c_void_p),
                ("m_q", C.c_void_p), ("contig", C.c_int32), ("global_k", C.c_int32)]


class _SplitResult(C.Structure):
    _fields_ = [("split", C.c_void_p), ("val_cluster", C.c_void_p), ("sp_off", C.c_void_p), ("sp", C.c_void_p), ("sp_val", C.c_void_p), ("sp_n0", C.c_void_p),
                ("piece_cap", C.c_uint64), ("n_pieces", C.c_uint64)]


class _AlignmentGroups(C.Structure):
    _fields_ = [("n_reads", C.c_int32)] + [(k, C.c_void_p) for k in ["grp_off", "seg_off", "upd_off", "update_at", "value", "n0", "n1", "nm", "nmm", "ndel", "nins", "strand"]] + \
               [("bypass_clustering", C.c_int32), ("read_type", C.c_int32), ("global_k", C.c_int32)]


class _MapqResult(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ["flag", "typeofaln", "issec", "supp", "mapq", "g_issec", "g_value", "g_n0", "g_n1", "g_nm", "order"]]


class _Refined(C.Structure):
    _fields_ = [("status", C.c_void_p), ("chrom", C.c_void_p), ("diag", C.c_void_p), ("r_off", C.c_void_p), ("r_q", C.c_void_p),
                ("r_t", C.c_void_p), ("r_tup", C.c_void_p), ("anchor_cap", C.c_uint64), ("n_anchors", C.c_uint64), ("rbox", C.c_void_p),
                ("eff", C.c_void_p), ("m_q_out", C.c_void_p), ("m_t_out", C.c_void_p), ("box_out", C.c_void_p), ("n_units", C.c_uint64),
                ("n_tasks", C.c_uint64)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("jobs", C.c_uint64), ("cells", C.c_uint64),
                ("algo_bytes", C.c_uint64)]


def library_path():
    return os.environ.get("LRA_B200_LIB") or os.path.join(_HERE, "liblra_b200.so")


def load_library():
    """Load liblra_b200.so (built in-tree by lra_b200/build.py).  Fails loudly if it is missing: no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LraB200Error(ECUDA, "CUDA library %s is missing; run `python -m lra_b200.build` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    L.lra_b200_version.restype = C.c_int
    L.lra_b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.lra_b200_destroy.argtypes = [C.c_void_p]
    L.lra_b200_destroy.restype = None
    L.lra_b200_last_error.argtypes = [C.c_void_p]
    L.lra_b200_last_error.restype = C.c_char_p
    L.lra_b200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_synchronize.argtypes = [C.c_void_p]
    L.lra_b200_seq_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.lra_b200_seq_from_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.lra_b200_seq_reupload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.lra_b200_seq_free.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_seq_free.restype = None
    L.lra_b200_seq_length.argtypes = [C.c_void_p]
    L.lra_b200_seq_length.restype = C.c_uint64
    L.lra_b200_seq_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lra_b200_aog_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_AogJobs), C.POINTER(_AogResult)]
    L.lra_b200_aog_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_AogJobs), C.POINTER(_AogResult)]
    L.lra_b200_indel_dp_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_IrGroups), C.POINTER(_IrResult)]
    L.lra_b200_indel_dp_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_IrGroups), C.POINTER(_IrResult)]
    L.lra_b200_indel_refine_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_IrSegments), C.POINTER(_IrSegResult)]
    L.lra_b200_indel_refine_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_IrSegments), C.POINTER(_IrSegResult)]
    L.lra_b200_index_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.lra_b200_index_free.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_index_free.restype = None
    L.lra_b200_seq_revcomp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.lra_b200_seed_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_SeedReads), C.POINTER(_SeedResult)]
    L.lra_b200_calc_stats_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_IrSegments), C.c_void_p, C.POINTER(_StatsResult)]
    L.lra_b200_sort_matches_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.lra_b200_global_chain_batch.argtypes = [C.c_void_p] * 3 + [C.c_int32] + [C.c_void_p] * 4
    L.lra_b200_refine_breakpoint_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Breakpoints), C.POINTER(_BreakpointResult)]
    L.lra_b200_linear_extend_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_ExtendParts), C.POINTER(_Extended)]
    L.lra_b200_split_chains_batch.argtypes = [C.c_void_p, C.POINTER(_AnchorChains), C.POINTER(_SplitChainsOut)]
    L.lra_b200_refine_space_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Spaces), C.POINTER(_SpaceResult)]
    L.lra_b200_refine_linear_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_LinearGaps), C.POINTER(_AogResult)]
    L.lra_b200_switch_to_original_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.lra_b200_split_rough_batch.argtypes = [C.c_void_p, C.POINTER(_RoughLists), C.POINTER(_SplitRoughResult)]
    L.lra_b200_store_diagonal_batch.argtypes = [C.c_void_p, C.POINTER(_CleanedLists), C.POINTER(_DiagClusters)]
    L.lra_b200_trim_splitchains_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lra_b200_merge_chain_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.lra_b200_switchindex_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.lra_b200_linear_extend_chains_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_ExtendChains), C.POINTER(_ExtendedChains)]
    L.lra_b200_chain_filter_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.lra_b200_clean_off_diagonal_batch.argtypes = [C.c_void_p, C.POINTER(_AnchorLists), C.POINTER(_CleanOpts), C.POINTER(_CleanResult)]
    L.lra_b200_split_clusters_batch.argtypes = [C.c_void_p, C.POINTER(_ReadClusters), C.POINTER(_SplitResult)]
    L.lra_b200_mapq_batch.argtypes = [C.c_void_p, C.POINTER(_AlignmentGroups), C.POINTER(_MapqResult)]
    L.lra_b200_lindex_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_void_p)]
    L.lra_b200_lindex_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.POINTER(C.c_void_p)]
    L.lra_b200_lindex_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.lra_b200_lindex_sizes.restype = None
    L.lra_b200_lindex_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lra_b200_lindex_free.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_lindex_free.restype = None
    L.lra_b200_refine_clusters_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Clusters), C.POINTER(_Refined)]
    L.lra_b200_refine_clusters_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Clusters), C.c_uint64, C.POINTER(_Refined)]
    L.lra_b200_refine_splitchains_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_SplitChains), C.POINTER(_Refined)]
    L.lra_b200_refine_splitchains_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_SplitChains), C.c_uint64, C.POINTER(_Refined)]
    L.lra_b200_calc_stats_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_IrSegments), C.c_void_p, C.POINTER(_StatsResult)]
    L.lra_b200_sdp_batch.argtypes = [C.c_void_p, C.POINTER(_SdpProblems), C.POINTER(_SdpResult)]
    L.lra_b200_init_pwl.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lra_b200_map_opts_preset.argtypes = [C.c_char_p, C.POINTER(MapOpts)]
    L.lra_b200_format_sam.restype = C.c_int64
    L.lra_b200_format_sam.argtypes = [C.POINTER(MapOpts), C.POINTER(_MapResult), C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_int64]
    L.lra_b200_mapper_create.argtypes = [C.c_void_p, C.POINTER(MapOpts), C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.lra_b200_mapper_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_mapper_destroy.restype = None
    L.lra_b200_map_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(_MapResult)]
    L.lra_b200_gindex_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.lra_b200_index_size.argtypes = [C.c_void_p]
    L.lra_b200_index_size.restype = C.c_uint64
    L.lra_b200_index_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lra_b200_index_free.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_index_free.restype = None
    L.lra_b200_readset_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.lra_b200_readset_free.argtypes = [C.c_void_p, C.c_void_p]
    L.lra_b200_readset_free.restype = None
    L.lra_b200_map_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.lra_b200_map_download.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_MapResult)]
    L.lra_b200_last_kernel_stats.argtypes = [C.c_void_p, C.POINTER(KernelStat), C.c_int]
    L.lra_b200_launch_count.argtypes = [C.c_void_p]
    L.lra_b200_launch_count.restype = C.c_uint64
    _LIB = L
    return L


def init_pwl(gapopen, gapextend, gaproot, ceil1, ceil2):
    """InitPWL (SubRountine.h:43-101) on the host: (stops int64[25], slope f32[25], inter f32[25], ceil1, ceil2)."""
    stops = np.zeros(25, np.int64); slope = np.zeros(25, np.float32); inter = np.zeros(25, np.float32)
    load_library().lra_b200_init_pwl(gapopen, gapextend, gaproot, ceil1, ceil2, stops.ctypes.data, slope.ctypes.data, inter.ctypes.data)
    return stops, slope, inter, ceil1, ceil2


def map_opts_preset(mode):
    """The align preset of `lra align -ONT | -CLR` (lra.cpp:339-431)."""
    o = MapOpts()
    rc = load_library().lra_b200_map_opts_preset(mode.encode(), C.byref(o))
    if rc:
        raise LraB200Error(rc, "unknown / unsupported preset %r" % mode)
    return o


def read_mms(path):
    """<ref>.mms (MMIndex.h:402-424): dict(k, names, pos (cumulative contig offsets), t (uint64 tuples), pos_t (uint32 positions))."""
    d = open(path, "rb").read()
    n = int(np.frombuffer(d, np.int64, 1, 0)[0]); k = int(np.frombuffer(d, np.int32, 1, 8)[0])
    nc = int(np.frombuffer(d, np.int32, 1, 12)[0]); o = 16
    names = []
    for _ in range(nc):
        ln = int(np.frombuffer(d, np.int32, 1, o)[0]); o += 4
        names.append(d[o:o + ln].decode()); o += ln
    hdr = np.frombuffer(d, np.uint64, nc + 1, o).copy(); o += 8 * (nc + 1)
    rec = np.frombuffer(d, np.dtype([("t", "<u8"), ("pos", "<u4"), ("pad", "<u4")]), n, o)
    return dict(k=k, names=names, hdr=hdr, t=np.ascontiguousarray(rec["t"]), pos=np.ascontiguousarray(rec["pos"]))


def _map_result_struct(res):
    return _MapResult(_ptr(res["status"]), _ptr(res["n_aln"]), _ptr(res["aln_nseg"]), _ptr(res["aln_seg0"]), _ptr(res["aln_rank"]), res["records"].ctypes.data,
                      len(res["records"]), int(res["n_records"]), _ptr(res["cigar"]), len(res["cigar"]), int(res["n_cigar"]), int(res.get("aligned_bases", 0)))


def format_sam(opts, res, names, reads, read_off, read_len, contig_names, runtime=0):
    """SAM records of a batch (lra_b200_format_sam): Alignment::PrintSAM for every printed segment, SimplePrintSAM for unaligned reads."""
    L = load_library()
    ms = _map_result_struct(res)
    nm = b"".join(n.encode() + b"\0" for n in names); cn = b"".join(n.encode() + b"\0" for n in contig_names)
    reads = np.ascontiguousarray(reads, np.uint8); ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
    need = -L.lra_b200_format_sam(C.byref(opts), C.byref(ms), len(names), nm, reads.ctypes.data, ro.ctypes.data, rl.ctypes.data, cn, len(contig_names), runtime, None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    n = L.lra_b200_format_sam(C.byref(opts), C.byref(ms), len(names), nm, reads.ctypes.data, ro.ctypes.data, rl.ctypes.data, cn, len(contig_names), runtime, buf, need + 16)
    return buf.raw[:n].decode()


def format_records_ref(opts, res, names, reads, read_off, read_len, contig_names, genome, contig_off, fmt="a", print_md=False, runtime=0):
    """lra_b200_format_records_ref: the printers that read the reference bases -- fmt 'a' (PrintPairwise, `-p a`), or fmt 's' with print_md (MD:Z:, `--printMD`)."""
    L = load_library()
    L.lra_b200_format_records_ref.restype = C.c_int64
    L.lra_b200_format_records_ref.argtypes = [C.POINTER(MapOpts), C.POINTER(_MapResult), C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
    ms = _map_result_struct(res)
    nm = names_blob(names); cn = names_blob(contig_names)
    reads = np.ascontiguousarray(reads, np.uint8); ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
    g = np.ascontiguousarray(genome, np.uint8); co = np.ascontiguousarray(contig_off, np.uint64)
    clen = np.diff(co).astype(np.uint64)
    args = (C.byref(opts), C.byref(ms), len(names), nm, reads.ctypes.data, None, ro.ctypes.data, rl.ctypes.data, cn, clen.ctypes.data, len(contig_names), g.ctypes.data, co.ctypes.data,
            ord(fmt), 1 if print_md else 0, runtime)
    need = -L.lra_b200_format_records_ref(*args, None, 0)
    if need <= 0:
        return ""
    buf = C.create_string_buffer(int(need) + 16)
    n = L.lra_b200_format_records_ref(*args, buf, need + 16)
    return buf.raw[:n].decode()


def names_blob(names):
    """n NUL-terminated strings back to back (the `names` / `contig_names` argument of lra_b200_format_sam)."""
    return b"".join(n.encode() + b"\0" for n in names)


def format_sam_into(opts, res, names_b, reads, read_off, read_len, contig_b, n_contigs, out, runtime=0):
    """lra_b200_format_sam into a caller-owned uint8 buffer (one pass); returns the number of bytes written."""
    L = load_library()
    ms = _map_result_struct(res)
    reads = np.ascontiguousarray(reads, np.uint8); ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
    n = L.lra_b200_format_sam(C.byref(opts), C.byref(ms), len(rl), names_b, reads.ctypes.data, ro.ctypes.data, rl.ctypes.data, contig_b, n_contigs, runtime, out.ctypes.data, len(out))
    if n < 0:
        raise LraB200Error(EOVERFLOW, "format_sam: buffer of %d bytes too small, %d needed" % (len(out), -n))
    return int(n)


_PART_KEYS = ("status", "n_aln", "aln_nseg", "aln_seg0", "aln_rank", "records", "cigar")


def result_from_parts(parts):
    """Inverse of Mapper.download_device: 1-D CPU tensors (per-read arrays, record bytes, cigar words, [n_records, n_cigar, aligned_bases]) -> the result dict."""
    res = {}
    for k, t in zip(_PART_KEYS, parts):
        a = t.numpy()
        res[k] = a.view(RECORD) if k == "records" else a
    meta = parts[len(_PART_KEYS)].numpy()
    res["n_records"] = int(meta[0]); res["n_cigar"] = int(meta[1]); res["aligned_bases"] = int(meta[2])
    return res


class Mapper:
    """lra_b200_mapper: the MapRead seam for one reference (genome + <ref>.mms + <ref>.gli) and one preset."""

    def __init__(self, ctx, opts, genome, hdr, mms, gli=None):
        self.ctx = ctx; self.opts = opts
        g = np.ascontiguousarray(genome, np.uint8); hdr = np.ascontiguousarray(hdr, np.uint64)
        h = C.c_void_p()
        if gli is not None:
            so = np.ascontiguousarray(gli["seq_offsets"], np.uint64); tb = np.ascontiguousarray(gli["tuple_boundaries"], np.uint64); mn = np.ascontiguousarray(gli["minimizers"], np.uint32)
            args = (_ptr(so), _ptr(tb), len(so), _ptr(mn), len(mn))
        else:
            args = (None, None, 0, None, 0)
        t = np.ascontiguousarray(mms["t"], np.uint64); p = np.ascontiguousarray(mms["pos"], np.uint32)
        ctx._check(ctx.lib.lra_b200_mapper_create(ctx.h, C.byref(opts), _ptr(g), len(g), _ptr(hdr), len(hdr) - 1, _ptr(t), _ptr(p), len(t), *args, C.byref(h)))
        self.h = h

    @staticmethod
    def _result_buffers(n, bases, record_cap=None, cigar_cap=None):
        record_cap = record_cap or 3 * n + 1024 + 65536; cigar_cap = cigar_cap or int(bases) + 64 * n + 4096
        return dict(status=np.zeros(max(n, 1), np.int32), n_aln=np.zeros(max(n, 1), np.int32), aln_nseg=np.zeros(4 * max(n, 1), np.int32), aln_seg0=np.zeros(4 * max(n, 1), np.int32),
                    aln_rank=np.zeros(4 * max(n, 1), np.int32), records=np.zeros(record_cap, RECORD), n_records=0, cigar=np.zeros(cigar_cap, np.uint32), n_cigar=0)

    def map_batch(self, reads, read_off, read_len, record_cap=None, cigar_cap=None, res=None):
        """MapRead over a batch of reads given as host buffers (H2D of the reads and D2H of the records inside the call)."""
        reads = np.ascontiguousarray(reads, np.uint8); ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
        n = len(rl)
        res = res if res is not None else self._result_buffers(n, rl.sum(), record_cap, cigar_cap)
        ms = _map_result_struct(res)
        self.ctx._check(self.ctx.lib.lra_b200_map_batch(self.ctx.h, self.h, _ptr(reads), len(reads), _ptr(ro), _ptr(rl), n, C.byref(ms)))
        res["n_records"] = int(ms.n_records); res["n_cigar"] = int(ms.n_cigar); res["aligned_bases"] = int(ms.aligned_bases)
        return res

    def upload(self, reads, read_off, read_len):
        """lra_b200_readset_upload: the batch packed and resident in HBM; returns an opaque handle (free with free_readset)."""
        reads = np.ascontiguousarray(reads, np.uint8); ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.lra_b200_readset_upload(self.ctx.h, _ptr(reads), len(reads), _ptr(ro), _ptr(rl), len(rl), C.byref(h)))
        return h

    def upload_device(self, reads_t, read_off, read_len):
        """lra_b200_readset_upload from a uint8 CUDA tensor (a shard received over NCCL): no host bounce of the bases."""
        ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.lra_b200_readset_upload(self.ctx.h, C.c_void_p(reads_t.data_ptr()), int(reads_t.numel()), _ptr(ro), _ptr(rl), len(rl), C.byref(h)))
        return h

    def download_device(self, n, bases, device):
        """lra_b200_map_download into CUDA tensors (to be sent over NCCL): [status, n_aln, aln_nseg, aln_seg0, aln_rank, record bytes, cigar, meta]."""
        import torch
        record_cap = 3 * n + 1024; cigar_cap = int(bases) + 64 * n + 4096
        t = dict(status=torch.empty(max(n, 1), dtype=torch.int32, device=device), n_aln=torch.empty(max(n, 1), dtype=torch.int32, device=device),
                 aln_nseg=torch.empty(4 * max(n, 1), dtype=torch.int32, device=device), aln_seg0=torch.empty(4 * max(n, 1), dtype=torch.int32, device=device),
                 aln_rank=torch.empty(4 * max(n, 1), dtype=torch.int32, device=device), records=torch.empty(record_cap * RECORD.itemsize, dtype=torch.uint8, device=device),
                 cigar=torch.empty(cigar_cap, dtype=torch.int32, device=device))
        ms = _MapResult(*(C.c_void_p(t[k].data_ptr()) for k in ("status", "n_aln", "aln_nseg", "aln_seg0", "aln_rank")), t["records"].data_ptr(), record_cap, 0,
                        C.c_void_p(t["cigar"].data_ptr()), cigar_cap, 0, 0)
        self.ctx._check(self.ctx.lib.lra_b200_map_download(self.ctx.h, self.h, C.byref(ms)))
        nr, nc = int(ms.n_records), int(ms.n_cigar)
        meta = torch.tensor([nr, nc, int(ms.aligned_bases)], dtype=torch.int64, device=device)
        return [t["status"][:n], t["n_aln"][:n], t["aln_nseg"][:4 * n], t["aln_seg0"][:4 * n], t["aln_rank"][:4 * n], t["records"][:nr * RECORD.itemsize], t["cigar"][:nc].view(torch.int32), meta]

    def free_readset(self, h):
        self.ctx.lib.lra_b200_readset_free(self.ctx.h, h)

    def map_resident(self, readset):
        self.ctx._check(self.ctx.lib.lra_b200_map_resident(self.ctx.h, self.h, readset))

    def map_download(self, n, bases, res=None):
        res = res if res is not None else self._result_buffers(n, bases)
        ms = _map_result_struct(res)
        self.ctx._check(self.ctx.lib.lra_b200_map_download(self.ctx.h, self.h, C.byref(ms)))
        res["n_records"] = int(ms.n_records); res["n_cigar"] = int(ms.n_cigar); res["aligned_bases"] = int(ms.aligned_bases)
        return res

    def close(self):
        if self.h:
            self.ctx.lib.lra_b200_mapper_destroy(self.ctx.h, self.h); self.h = None


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class SeqArena:
    """Device-resident packed sequence arena (lra_b200_seq)."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def __len__(self):
        return int(self.ctx.lib.lra_b200_seq_length(self.handle))

    def reupload(self, ascii_bytes):
        buf = np.frombuffer(ascii_bytes, dtype=np.uint8) if not isinstance(ascii_bytes, np.ndarray) else ascii_bytes
        self.ctx._check(self.ctx.lib.lra_b200_seq_reupload(self.ctx.h, self.handle, _ptr(buf), len(buf)))

    def download(self):
        n = len(self)
        b2 = np.zeros((n + 15) // 16, np.uint32); nm = np.zeros((n + 31) // 32, np.uint32)
        self.ctx._check(self.ctx.lib.lra_b200_seq_download(self.ctx.h, self.handle, _ptr(b2), _ptr(nm)))
        return b2, nm

    def free(self):
        if self.handle:
            self.ctx.lib.lra_b200_seq_free(self.ctx.h, self.handle)
            self.handle = None


class LocalIndexImage:
    """A LocalIndex resident on the device (lra_b200_lindex)."""
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def sizes(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.lib.lra_b200_lindex_sizes(self.handle, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def download(self):
        """(win_off[n_win + 1], bnd[n_win + 1], mins) as in the reference's LocalIndex (seqOffsets without the leading 0 when the arena
        starts at 0; tupleBoundaries; minimizers as uint32)."""
        nw, nm = self.sizes()
        wo = np.zeros(nw + 1, np.uint64); bd = np.zeros(nw + 1, np.uint64); mn = np.zeros(max(nm, 1), np.uint32)
        self.ctx._check(self.ctx.lib.lra_b200_lindex_download(self.ctx.h, self.handle, _ptr(wo), _ptr(bd), _ptr(mn)))
        return wo, bd, mn[:nm]

    def free(self):
        if self.handle:
            self.ctx.lib.lra_b200_lindex_free(self.ctx.h, self.handle)
            self.handle = None


class Context:
    """One GPU context (lra_b200_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.lra_b200_create(C.byref(h), device)
        if rc != OK:
            raise LraB200Error(rc, self.lib.lra_b200_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.lra_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise LraB200Error(rc, self.lib.lra_b200_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.lra_b200_set_stream(self.h, cuda_stream_ptr))

    def synchronize(self):
        self._check(self.lib.lra_b200_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.lra_b200_launch_count(self.h))

    def seq_upload(self, ascii_bytes):
        buf = np.frombuffer(ascii_bytes, dtype=np.uint8) if not isinstance(ascii_bytes, np.ndarray) else ascii_bytes
        h = C.c_void_p()
        self._check(self.lib.lra_b200_seq_upload(self.h, _ptr(buf), len(buf), C.byref(h)))
        return SeqArena(self, h)

    def seq_from_device(self, dev_ptr, n):
        h = C.c_void_p()
        self._check(self.lib.lra_b200_seq_from_device(self.h, dev_ptr, n, C.byref(h)))
        return SeqArena(self, h)

    def kernel_stats(self):
        arr = (KernelStat * 32)()
        n = self.lib.lra_b200_last_kernel_stats(self.h, arr, 32)
        return [dict(name=arr[i].name.decode(), ms=float(arr[i].ms), jobs=int(arr[i].jobs), cells=int(arr[i].cells),
                     algo_bytes=int(arr[i].algo_bytes)) for i in range(min(n, 32))]

    # ---- a18
    def aog_batch(self, q, t, q_off, t_off, q_len, t_len, k, m, mm, indel, block_cap=None, out=None):
        """Host arrays in / host arrays out.  Returns dict(score, n_blocks, block_off, blocks[n,3], cells)."""
        n = len(q_off)
        q_off = np.ascontiguousarray(q_off, np.uint32); t_off = np.ascontiguousarray(t_off, np.uint32)
        q_len = np.ascontiguousarray(q_len, np.int32); t_len = np.ascontiguousarray(t_len, np.int32)
        k = np.ascontiguousarray(k, np.int32)
        if block_cap is None:
            block_cap = int(np.minimum(q_len, t_len).clip(min=0).sum()) + 1
        if out is None:
            out = dict(score=np.zeros(n, np.int32), n_blocks=np.zeros(n, np.int32), block_off=np.zeros(n, np.uint64),
                       blocks=np.zeros((max(1, block_cap), 3), np.uint32))
        jobs = _AogJobs(_ptr(q_off), _ptr(t_off), _ptr(q_len), _ptr(t_len), _ptr(k), n, m, mm, indel)
        res = _AogResult(_ptr(out["score"]), _ptr(out["n_blocks"]), _ptr(out["block_off"]), _ptr(out["blocks"]),
                         block_cap, 0, 0)
        rc = self.lib.lra_b200_aog_batch(self.h, q.handle, t.handle, C.byref(jobs), C.byref(res))
        out["n_blocks_total"] = int(res.n_blocks_total)
        out["cells"] = int(res.cells)
        self._last_needed = int(res.n_blocks_total)
        self._check(rc)
        return out

    def aog_batch_device(self, q, t, d_q_off, d_t_off, d_q_len, d_t_len, d_k, n, m, mm, indel, d_score, d_n_blocks,
                         d_block_off, d_blocks, block_cap):
        """Everything resident: arguments are raw device pointers (ints).  Returns (n_blocks_total, cells)."""
        jobs = _AogJobs(d_q_off, d_t_off, d_q_len, d_t_len, d_k, n, m, mm, indel)
        res = _AogResult(d_score, d_n_blocks, d_block_off, d_blocks, block_cap, 0, 0)
        self._check(self.lib.lra_b200_aog_batch_device(self.h, q.handle, t.handle, C.byref(jobs), C.byref(res)))
        return int(res.n_blocks_total), int(res.cells)


    # ---- a1-a5
    def index_upload(self, t, pos):
        t = np.ascontiguousarray(t, np.uint64); pos = np.ascontiguousarray(pos, np.uint32)
        h = C.c_void_p()
        self._check(self.lib.lra_b200_index_upload(self.h, _ptr(t), _ptr(pos), len(t), C.byref(h)))
        return h

    def index_free(self, h):
        self.lib.lra_b200_index_free(self.h, h)

    def seq_revcomp(self, reads, read_off, read_len, reuse=None):
        ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
        h = C.c_void_p(reuse.handle.value) if reuse is not None else C.c_void_p()
        if reuse is not None:
            reuse.handle = None
        self._check(self.lib.lra_b200_seq_revcomp(self.h, reads.handle, _ptr(ro), _ptr(rl), len(ro), C.byref(h)))
        return SeqArena(self, h)

    def seed_batch(self, reads, genome, index, read_off, read_len, k, w, max_freq, match_cap=None):
        """a2-a5 for a batch of reads.  Returns dict(match_off, q_t, q_pos, t_t, t_pos, strand, n_minimizers, n_matches)."""
        ro = np.ascontiguousarray(read_off, np.uint64); rl = np.ascontiguousarray(read_len, np.uint32)
        R = len(ro)
        cap = match_cap if match_cap is not None else int(rl.sum()) + 1024
        for _ in range(2):
            o = dict(match_off=np.zeros(R + 1, np.uint64), q_t=np.zeros(cap, np.uint64), q_pos=np.zeros(cap, np.uint32), t_t=np.zeros(cap, np.uint64),
                     t_pos=np.zeros(cap, np.uint32), strand=np.zeros(cap, np.uint8), n_minimizers=np.zeros(R, np.uint32))
            rd = _SeedReads(_ptr(ro), _ptr(rl), R, k, w, max_freq)
            res = _SeedResult(_ptr(o["match_off"]), _ptr(o["q_t"]), _ptr(o["q_pos"]), _ptr(o["t_t"]), _ptr(o["t_pos"]), _ptr(o["strand"]), cap, 0,
                              _ptr(o["n_minimizers"]))
            rc = self.lib.lra_b200_seed_batch(self.h, reads.handle, genome.handle, index, C.byref(rd), C.byref(res))
            o["n_matches"] = int(res.n_matches)
            if rc == EOVERFLOW and match_cap is None:
                cap = int(res.n_matches) + 16
                continue
            self._check(rc)
            return o
        self._check(rc)

    # ---- a6
    def sort_matches_batch(self, mode, q, t, seg_off, want_perm=True):
        """DiagonalSort (mode 0) / AntiDiagonalSort (1) / CartesianSort (2) / CartesianTargetSort (3) of every segment.  Returns (q, t, perm)."""
        q = np.array(q, np.uint32); t = np.array(t, np.uint32); so = np.ascontiguousarray(seg_off, np.uint64)
        perm = np.zeros(max(len(q), 1), np.uint32) if want_perm else None
        self._check(self.lib.lra_b200_sort_matches_batch(self.h, mode, _ptr(q) if len(q) else None, _ptr(t) if len(t) else None, _ptr(so), len(so) - 1,
                                                         _ptr(perm) if want_perm else None))
        return q, t, (perm[:len(q)] if want_perm else None)

    # ---- a7
    def clean_off_diagonal_batch(self, q, t, qt, list_off, strand, opts, hdr_pos):
        """CleanOffDiagonal of every anchor list (opts: dict of the Options fields, see include/lra_b200.h).  Returns dict(keep, freq, cnt, cl[N,7],
        cl_freq, n_cl)."""
        q = np.ascontiguousarray(q, np.uint32); t = np.ascontiguousarray(t, np.uint32); qt = np.ascontiguousarray(qt, np.uint64)
        lo = np.ascontiguousarray(list_off, np.uint64); st = np.ascontiguousarray(strand, np.uint8); hdr = np.ascontiguousarray(hdr_pos, np.uint64)
        N, n = len(q), len(lo) - 1
        o = dict(keep=np.zeros(max(N, 1), np.uint8), freq=np.zeros(max(N, 1), np.float32), cnt=np.zeros(max(N, 1), np.int32), cl=np.zeros((max(N, 1), 7), np.int32),
                 cl_freq=np.zeros(max(N, 1), np.float32), n_cl=np.zeros(max(n, 1), np.int32))
        al = _AnchorLists(n, _ptr(q) if N else None, _ptr(t) if N else None, _ptr(qt) if N else None, _ptr(lo), _ptr(st) if n else None, _ptr(hdr), len(hdr))
        op = _CleanOpts(*[int(opts[k]) for k, _ in _CleanOpts._fields_])
        r = _CleanResult(_ptr(o["keep"]), _ptr(o["freq"]), _ptr(o["cnt"]), _ptr(o["cl"]), _ptr(o["cl_freq"]), _ptr(o["n_cl"]))
        self._check(self.lib.lra_b200_clean_off_diagonal_batch(self.h, C.byref(al), C.byref(op), C.byref(r)))
        return {k: (v[:n] if k == "n_cl" else v[:N]) for k, v in o.items()}

    # ---- a9
    def split_clusters_batch(self, cl_off, box, strand, freq, m_off, m_q, contig, global_k, piece_cap=None):
        """SplitClusters + DecideSplitClustersValue for every read.  Returns dict(split, val_cluster, sp_off, sp[k,6], sp_val, sp_n0, n_pieces)."""
        co = np.ascontiguousarray(cl_off, np.uint64); bx = np.ascontiguousarray(box, np.uint32).reshape(-1); st = np.ascontiguousarray(strand, np.uint8)
        fr = np.ascontiguousarray(freq, np.float32); mo = np.ascontiguousarray(m_off, np.uint64); mq = np.ascontiguousarray(m_q, np.uint32)
        R, Cn = len(co) - 1, len(st)
        cap = piece_cap if piece_cap is not None else 8 * Cn + 64
        for _ in range(2):
            o = dict(split=np.zeros(max(Cn, 1), np.uint8), val_cluster=np.zeros(max(Cn, 1), np.int32), sp_off=np.zeros(R + 1, np.uint64), sp=np.zeros((cap, 6), np.uint32),
                     sp_val=np.zeros(cap, np.int32), sp_n0=np.zeros(cap, np.int32))
            rcs = _ReadClusters(R, _ptr(co), _ptr(bx) if Cn else None, _ptr(st) if Cn else None, _ptr(fr) if Cn else None, _ptr(mo), _ptr(mq) if len(mq) else None, contig, global_k)
            res = _SplitResult(_ptr(o["split"]), _ptr(o["val_cluster"]), _ptr(o["sp_off"]), _ptr(o["sp"]), _ptr(o["sp_val"]), _ptr(o["sp_n0"]), cap, 0)
            rc = self.lib.lra_b200_split_clusters_batch(self.h, C.byref(rcs), C.byref(res))
            o["n_pieces"] = int(res.n_pieces)
            if rc == EOVERFLOW and piece_cap is None:
                cap = int(res.n_pieces) + 16
                continue
            self._check(rc)
            n = o["n_pieces"]
            o["split"] = o["split"][:Cn]; o["val_cluster"] = o["val_cluster"][:Cn]; o["sp"] = o["sp"][:n]; o["sp_val"] = o["sp_val"][:n]; o["sp_n0"] = o["sp_n0"][:n]
            return o
        self._check(rc)

    # ---- a16
    def chain_filter_batch(self, mode, q, t, length, strand, chain_off):
        """The keep mask of every chain under one of the reference's chain filters (modes in include/lra_b200.h)."""
        q = np.ascontiguousarray(q, np.uint32); t = np.ascontiguousarray(t, np.uint32); ln = np.ascontiguousarray(length, np.uint32)
        st = np.ascontiguousarray(strand, np.uint8); co = np.ascontiguousarray(chain_off, np.uint64)
        keep = np.zeros(max(len(q), 1), np.uint8)
        n = len(q)
        self._check(self.lib.lra_b200_chain_filter_batch(self.h, mode, _ptr(q) if n else None, _ptr(t) if n else None, _ptr(ln) if n else None, _ptr(st) if n else None,
                                                         _ptr(co), len(co) - 1, _ptr(keep)))
        return keep[:n]

    # ---- a20
    def refine_breakpoint_batch(self, reads_fwd, reads_rc, genome, bp):
        """RefineBreakpoint for every pair (bp: dict(lf, ll, rf, rl [n,3]; lstrand, rstrand, read_off, read_len, lchrom_off, rchrom_off,
        lchrom_len, rchrom_len)).  Returns dict(mode[n,2], n_out[n,2], bound[n,2,3], out[n,2,512,3], refined[n])."""
        n = len(bp["lstrand"])
        a = {k: np.ascontiguousarray(bp[k], np.uint32).reshape(-1) for k in ["lf", "ll", "rf", "rl", "read_len", "lchrom_len", "rchrom_len"]}
        a.update({k: np.ascontiguousarray(bp[k], np.uint8) for k in ["lstrand", "rstrand"]})
        a.update({k: np.ascontiguousarray(bp[k], np.uint64) for k in ["read_off", "lchrom_off", "rchrom_off"]})
        o = dict(mode=np.zeros((max(n, 1), 2), np.int32), n_out=np.zeros((max(n, 1), 2), np.int32), bound=np.zeros((max(n, 1), 2, 3), np.uint32),
                 out=np.zeros((max(n, 1), 2, 512, 3), np.uint32), refined=np.zeros(max(n, 1), np.int32))
        b = _Breakpoints(n, _ptr(a["lf"]), _ptr(a["ll"]), _ptr(a["rf"]), _ptr(a["rl"]), _ptr(a["lstrand"]), _ptr(a["rstrand"]), _ptr(a["read_off"]), _ptr(a["read_len"]),
                         _ptr(a["lchrom_off"]), _ptr(a["rchrom_off"]), _ptr(a["lchrom_len"]), _ptr(a["rchrom_len"]))
        r = _BreakpointResult(_ptr(o["mode"]), _ptr(o["n_out"]), _ptr(o["bound"]), _ptr(o["out"]), _ptr(o["refined"]))
        self._check(self.lib.lra_b200_refine_breakpoint_batch(self.h, reads_fwd.handle, reads_rc.handle, genome.handle, C.byref(b), C.byref(r)))
        return {k: v[:n] for k, v in o.items()}

    # ---- a15
    def linear_extend_batch(self, reads, genome, ep, K, skipsorting, trim):
        """LinearExtend (GenomePairs overload) + DecideCoordinates [+ TrimOverlappedAnchors] for every group of parts (ep: dict(g_off, p_off, p_strand,
        chrom_off, chrom_len, read_off, read_len, q, t)).  Returns dict(e_off, q, t, len, box[g,4])."""
        a = {k: np.ascontiguousarray(ep[k], np.uint64) for k in ["g_off", "p_off", "chrom_off", "read_off"]}
        a.update({k: np.ascontiguousarray(ep[k], np.uint32) for k in ["chrom_len", "read_len", "q", "t"]})
        a["p_strand"] = np.ascontiguousarray(ep["p_strand"], np.uint8)
        G = len(a["g_off"]) - 1; N = len(a["q"])
        o = dict(e_off=np.zeros(G + 1, np.uint64), q=np.zeros(max(N, 1), np.uint32), t=np.zeros(max(N, 1), np.uint32), len=np.zeros(max(N, 1), np.int32),
                 box=np.zeros((max(G, 1), 4), np.uint32))
        p = lambda x: _ptr(x) if x.size else None
        e = _ExtendParts(G, _ptr(a["g_off"]), p(a["p_off"]), p(a["p_strand"]), p(a["chrom_off"]), p(a["chrom_len"]), p(a["read_off"]), p(a["read_len"]), p(a["q"]), p(a["t"]),
                         K, int(skipsorting), int(trim))
        r = _Extended(_ptr(o["e_off"]), _ptr(o["q"]), _ptr(o["t"]), _ptr(o["len"]), max(N, 1), 0, _ptr(o["box"]))
        self._check(self.lib.lra_b200_linear_extend_batch(self.h, reads.handle, genome.handle, C.byref(e), C.byref(r)))
        n = int(r.n_total)
        for k in ("q", "t", "len"):
            o[k] = o[k][:n]
        o["box"] = o["box"][:G]
        return o

    def linear_extend_chains_batch(self, reads, genome, cd, K, skiprepetitive=1, trim=1, merge_dist=100):
        """LinearExtend_chain (the high-accuracy overload) + MergeMatchesSameDiag for every chain (cd: dict(ch_off, ch, cl_off, q, t, box[n,4], strand, freq,
        chrom_off, chrom_len, read_off, read_len)).  Returns dict(e_off, q, t, len, ovp, md_head, box[u,4], overlap[u])."""
        a = {k: np.ascontiguousarray(cd[k], np.uint64) for k in ["ch_off", "cl_off", "chrom_off", "read_off"]}
        a.update({k: np.ascontiguousarray(cd[k], np.uint32).reshape(-1) for k in ["ch", "chrom_len", "read_len", "q", "t", "box"]})
        a["strand"] = np.ascontiguousarray(cd["strand"], np.uint8); a["freq"] = np.ascontiguousarray(cd["freq"], np.float32)
        NC = len(a["ch_off"]) - 1; CL = len(a["cl_off"]) - 1; U = len(a["ch"])
        sizes = np.diff(a["cl_off"].astype(np.int64))
        cap = int(sizes[a["ch"]].sum()) if U else 0
        o = dict(e_off=np.zeros(U + 1, np.uint64), q=np.zeros(max(cap, 1), np.uint32), t=np.zeros(max(cap, 1), np.uint32), len=np.zeros(max(cap, 1), np.int32),
                 ovp=np.zeros(max(cap, 1), np.uint8), md_head=np.zeros(max(cap, 1), np.uint8), box=np.zeros((max(U, 1), 4), np.uint32), overlap=np.zeros(max(U, 1), np.int32))
        p = lambda x: _ptr(x) if x.size else None
        e = _ExtendChains(NC, _ptr(a["ch_off"]), p(a["ch"]), CL, _ptr(a["cl_off"]), p(a["q"]), p(a["t"]), p(a["box"]), p(a["strand"]), p(a["freq"]), p(a["chrom_off"]),
                          p(a["chrom_len"]), p(a["read_off"]), p(a["read_len"]), K, int(skiprepetitive), int(trim), int(merge_dist))
        r = _ExtendedChains(_ptr(o["e_off"]), _ptr(o["q"]), _ptr(o["t"]), _ptr(o["len"]), _ptr(o["ovp"]), _ptr(o["md_head"]), max(cap, 1), 0, _ptr(o["box"]), _ptr(o["overlap"]))
        self._check(self.lib.lra_b200_linear_extend_chains_batch(self.h, reads.handle, genome.handle, C.byref(e), C.byref(r)))
        n = int(r.n_total)
        for k in ("q", "t", "len", "ovp", "md_head"):
            o[k] = o[k][:n]
        o["box"] = o["box"][:U]; o["overlap"] = o["overlap"][:U]
        return o

    # ---- a11
    def split_chains_batch(self, ac, hdr_pos, splitdist=50000, bypass=0):
        """SPLITChain(UltimateChain) + MergeSplitchainINS + RemoveSpuriousSplitChain for every chain (ac: dict(c_off, q, t, len, strand, cnum, link -- link has
        one entry per anchor, the last of each chain unused)).  Returns the slot-layout arrays of lra_b200_split_chain_result; see split_chain_view."""
        co = np.ascontiguousarray(ac["c_off"], np.uint64); NC = len(co) - 1; N = int(co[-1])
        a = dict(q=np.ascontiguousarray(ac["q"], np.uint32), t=np.ascontiguousarray(ac["t"], np.uint32), len=np.ascontiguousarray(ac["len"], np.int32),
                 strand=np.ascontiguousarray(ac["strand"], np.uint8), cnum=np.ascontiguousarray(ac["cnum"], np.int32), link=np.ascontiguousarray(ac["link"], np.uint8))
        assert all(len(v) == N for v in a.values())
        hdr = np.ascontiguousarray(hdr_pos, np.uint64)
        Np = max(N, 1)
        o = dict(n_sp=np.zeros(max(NC, 1), np.int32), n_link=np.zeros(max(NC, 1), np.int32), sp_off=np.zeros(Np + NC + 1, np.int32), ci_off=np.zeros(Np + NC + 1, np.int32),
                 sptc=np.zeros(Np, np.int32), ci=np.zeros(Np, np.int32), sp_lk=np.zeros(Np, np.uint8), sp_box=np.zeros((Np, 4), np.uint32), sp_chrom=np.zeros(Np, np.int32),
                 sp_type=np.zeros(Np, np.uint8), sp_strand=np.zeros(Np, np.uint8), sp_link=np.zeros(Np, np.uint8))
        p = lambda x: _ptr(x) if x.size else None
        e = _AnchorChains(NC, _ptr(co), p(a["q"]), p(a["t"]), p(a["len"]), p(a["strand"]), p(a["cnum"]), p(a["link"]), _ptr(hdr), len(hdr), int(splitdist), int(bypass))
        r = _SplitChainsOut(*[_ptr(o[k]) for k in ["n_sp", "n_link", "sp_off", "ci_off", "sptc", "ci", "sp_lk", "sp_box", "sp_chrom", "sp_type", "sp_strand", "sp_link"]])
        self._check(self.lib.lra_b200_split_chains_batch(self.h, C.byref(e), C.byref(r)))
        return o

    def merge_chain_batch(self, sp, sc_off, chrom, strand, box):
        """MergeChain for every split chain; returns the head flags (1 = the entry starts a new Merge_SplitChain)."""
        sp = np.ascontiguousarray(sp, np.int32); so = np.ascontiguousarray(sc_off, np.uint64)
        chrom = np.ascontiguousarray(chrom, np.int32); strand = np.ascontiguousarray(strand, np.uint8); box = np.ascontiguousarray(box, np.uint32).reshape(-1)
        head = np.zeros(max(len(sp), 1), np.uint8)
        p = lambda x: _ptr(x) if x.size else None
        self._check(self.lib.lra_b200_merge_chain_batch(self.h, p(sp), _ptr(so), len(so) - 1, p(chrom), p(strand), p(box), len(chrom), _ptr(head)))
        return head[:len(sp)]

    def switchindex_batch(self, ch, link, c_off, coarse, cq):
        """switchindex for every chain (link: one byte per entry, the last of each chain unused).  Returns (ch, link, n_out, nl_out) in slot layout."""
        ch = np.array(ch, np.int32); link = np.array(link, np.uint8); co = np.ascontiguousarray(c_off, np.uint64)
        coarse = np.ascontiguousarray(coarse, np.int32); cq = np.ascontiguousarray(cq, np.uint32).reshape(-1)
        NC = len(co) - 1
        n_out = np.zeros(max(NC, 1), np.int32); nl_out = np.zeros(max(NC, 1), np.int32)
        p = lambda x: _ptr(x) if x.size else None
        self._check(self.lib.lra_b200_switchindex_batch(self.h, p(ch), p(link), _ptr(co), NC, p(coarse), len(coarse), p(cq), len(cq) // 2, _ptr(n_out), _ptr(nl_out)))
        return ch, link, n_out[:NC], nl_out[:NC]

    # ---- a17 (leaf)
    def refine_linear_batch(self, reads, genome, gaps, m, mm, indel, local_band, block_cap=None):
        """RefineByLinearAlignment for every gap (gaps: dict(cur_read_end, next_read_start, cur_genome_end, next_genome_start, read_off, chrom_off)).
        Returns dict(score, n_blocks, block_off, blocks[n,3]) with blocks in read / contig coordinates."""
        a = {k: np.ascontiguousarray(gaps[k], np.uint32) for k in ["cur_read_end", "next_read_start", "cur_genome_end", "next_genome_start", "read_off", "chrom_off"]}
        n = len(a["read_off"])
        if block_cap is None:
            ql = (a["next_read_start"] - a["cur_read_end"]).astype(np.int32).astype(np.int64); tl = (a["next_genome_start"] - a["cur_genome_end"]).astype(np.int32).astype(np.int64)
            block_cap = int(np.minimum(ql, tl).clip(min=0).sum()) + 1
        out = dict(score=np.zeros(max(n, 1), np.int32), n_blocks=np.zeros(max(n, 1), np.int32), block_off=np.zeros(max(n, 1), np.uint64), blocks=np.zeros((max(1, block_cap), 3), np.uint32))
        p = lambda x: _ptr(x) if x.size else None
        g = _LinearGaps(n, p(a["cur_read_end"]), p(a["next_read_start"]), p(a["cur_genome_end"]), p(a["next_genome_start"]), p(a["read_off"]), p(a["chrom_off"]), m, mm, indel, local_band)
        res = _AogResult(_ptr(out["score"]), _ptr(out["n_blocks"]), _ptr(out["block_off"]), _ptr(out["blocks"]), block_cap, 0, 0)
        self._check(self.lib.lra_b200_refine_linear_batch(self.h, reads.handle, genome.handle, C.byref(g), C.byref(res)))
        out["n_blocks_total"] = int(res.n_blocks_total)
        for k in ("score", "n_blocks", "block_off"):
            out[k] = out[k][:n]
        return out

    # ---- a14 (core, small spaces)
    def refine_space_batch(self, reads, genome, sp, K, m, mm, indel, W=10, local_max_freq=30):
        """RefineSpace for every space (sp: dict(qs, qe, ts, te, lrts, lrlength, read_off, read_len, chrom_off, flip[, diag = refineSpaceDiag per space, needed when a
        space has 1000 bases or more on an axis])).  Returns dict(pair_off, n_pairs, identity, pq, pt) in slot layout."""
        a = {k: np.ascontiguousarray(sp[k], np.uint32) for k in ["qs", "qe", "ts", "te", "lrts", "lrlength", "read_off", "read_len", "chrom_off"]}
        a["flip"] = np.ascontiguousarray(sp["flip"], np.uint8)
        diag = np.ascontiguousarray(sp["diag"], np.int32) if "diag" in sp else None
        n = len(a["qs"])
        ql = a["qe"].astype(np.int64) - a["qs"]; tl = a["te"].astype(np.int64) - a["ts"] + a["lrlength"]
        cap = int((np.minimum(ql, tl).clip(min=0) // K + 1).sum()) + 1
        p = lambda x: _ptr(x) if x.size else None
        for _ in range(2):
            o = dict(pair_off=np.zeros(n + 1, np.uint64), n_pairs=np.zeros(max(n, 1), np.int32), identity=np.zeros(max(n, 1), np.float32), pq=np.zeros(cap, np.uint32),
                     pt=np.zeros(cap, np.uint32))
            e = _Spaces(n, *[p(a[k]) for k in ["qs", "qe", "ts", "te", "lrts", "lrlength", "read_off", "read_len", "chrom_off", "flip"]], K, m, mm, indel, W, local_max_freq,
                        _ptr(diag) if diag is not None and diag.size else None)
            r = _SpaceResult(_ptr(o["pair_off"]), _ptr(o["n_pairs"]), _ptr(o["identity"]), _ptr(o["pq"]), _ptr(o["pt"]), cap, 0)
            rc = self.lib.lra_b200_refine_space_batch(self.h, reads.handle, genome.handle, C.byref(e), C.byref(r))
            if rc == EOVERFLOW:        # the pairs of the minimizer branch are only known after its count pass
                cap = int(r.n_pairs_total) + 1
                continue
            break
        self._check(rc)
        o["n_pairs"] = o["n_pairs"][:n]; o["identity"] = o["identity"][:n]
        return o

    def switch_to_original_batch(self, run_start, run_end, coarse):
        """SwitchToOriginalAnchors for every FinalChain entry.  Returns (off, chain, cluster_index)."""
        rs = np.ascontiguousarray(run_start, np.int32); re_ = np.ascontiguousarray(run_end, np.int32); co = np.ascontiguousarray(coarse, np.int32)
        n = len(rs); cap = int(np.clip(re_.astype(np.int64) - rs, 0, None).sum())
        off = np.zeros(n + 1, np.uint64); chain = np.zeros(max(cap, 1), np.uint32); ci = np.zeros(max(cap, 1), np.int32)
        tot = C.c_uint64(0)
        p = lambda x: _ptr(x) if x.size else None
        self._check(self.lib.lra_b200_switch_to_original_batch(self.h, p(rs), p(re_), p(co), n, _ptr(off), _ptr(chain), _ptr(ci), cap, C.byref(tot)))
        return off, chain[:int(tot.value)], ci[:int(tot.value)]

    # ---- a8 (first half)
    def split_rough_batch(self, rl, globalK, max_gap, min_cluster_size, max_diag):
        """SplitRoughClustersWithGaps for every anchor list (rl: dict(l_off, lr_off, q, t, r_start, r_end, r_box, r_strand, r_freq, r_chrom)).
        Returns the slot-layout arrays of lra_b200_split_rough_result (base of list l = l_off[l] + lr_off[l])."""
        lo = np.ascontiguousarray(rl["l_off"], np.uint64); lro = np.ascontiguousarray(rl["lr_off"], np.uint64)
        NL = len(lo) - 1; T = int(lo[-1]) + int(lro[-1]) + 1
        a = dict(q=np.ascontiguousarray(rl["q"], np.uint32), t=np.ascontiguousarray(rl["t"], np.uint32), r_start=np.ascontiguousarray(rl["r_start"], np.int32),
                 r_end=np.ascontiguousarray(rl["r_end"], np.int32), r_box=np.ascontiguousarray(rl["r_box"], np.uint32).reshape(-1), r_strand=np.ascontiguousarray(rl["r_strand"], np.uint8),
                 r_freq=np.ascontiguousarray(rl["r_freq"], np.float32), r_chrom=np.ascontiguousarray(rl["r_chrom"], np.int32))
        o = dict(n_split=np.zeros(max(NL, 1), np.int32), n_piece=np.zeros(max(NL, 1), np.int32), s_start=np.zeros(T, np.int32), s_end=np.zeros(T, np.int32),
                 s_coarse=np.zeros(T, np.int32), s_chrom=np.zeros(T, np.int32), s_box=np.zeros((T, 4), np.uint32), s_strand=np.zeros(T, np.uint8), s_freq=np.zeros(T, np.float32),
                 p_cluster=np.zeros(T, np.int32), p_start=np.zeros(T, np.int32), p_end=np.zeros(T, np.int32))
        p = lambda x: _ptr(x) if x.size else None
        e = _RoughLists(NL, _ptr(lo), _ptr(lro), p(a["q"]), p(a["t"]), p(a["r_start"]), p(a["r_end"]), p(a["r_box"]), p(a["r_strand"]), p(a["r_freq"]), p(a["r_chrom"]),
                        globalK, max_gap, min_cluster_size, max_diag)
        r = _SplitRoughResult(*[_ptr(o[k]) for k in ["n_split", "n_piece", "s_start", "s_end", "s_coarse", "s_chrom", "s_box", "s_strand", "s_freq", "p_cluster", "p_start", "p_end"]])
        self._check(self.lib.lra_b200_split_rough_batch(self.h, C.byref(e), C.byref(r)))
        return o

    def store_diagonal_batch(self, cl, hdr_pos, globalK, max_diag, min_cluster_size, min_cluster_length, bypass):
        """StoreDiagonalClusters for every cleaned anchor list (cl: dict(l_off, q, t, qt, freq, strand)).  Returns the slot-layout arrays of lra_b200_diag_clusters."""
        lo = np.ascontiguousarray(cl["l_off"], np.uint64); NL = len(lo) - 1; N = max(int(lo[-1]), 1)
        a = dict(q=np.ascontiguousarray(cl["q"], np.uint32), t=np.ascontiguousarray(cl["t"], np.uint32), qt=np.ascontiguousarray(cl["qt"], np.uint64),
                 freq=np.ascontiguousarray(cl["freq"], np.float32), strand=np.ascontiguousarray(cl["strand"], np.uint8))
        hdr = np.ascontiguousarray(hdr_pos, np.uint64)
        o = dict(n_cl=np.zeros(max(NL, 1), np.int32), c_start=np.zeros(N, np.int32), c_end=np.zeros(N, np.int32), c_chrom=np.zeros(N, np.int32), c_box=np.zeros((N, 4), np.uint32),
                 c_freq=np.zeros(N, np.float32))
        p = lambda x: _ptr(x) if x.size else None
        e = _CleanedLists(NL, _ptr(lo), p(a["q"]), p(a["t"]), p(a["qt"]), p(a["freq"]), p(a["strand"]), _ptr(hdr), len(hdr), globalK, max_diag, min_cluster_size,
                          min_cluster_length, int(bypass))
        r = _DiagClusters(*[_ptr(o[k]) for k in ["n_cl", "c_start", "c_end", "c_chrom", "c_box", "c_freq"]])
        self._check(self.lib.lra_b200_store_diagonal_batch(self.h, C.byref(e), C.byref(r)))
        return o

    def trim_splitchains_batch(self, cq, ct, c_off, strand, q, t, m_off):
        """TrimSplitChainDiagonal for every split chain.  Returns (q, t, keep, removed): the anchors in the reference's order and which of them stay."""
        cq = np.ascontiguousarray(cq, np.uint32); ct = np.ascontiguousarray(ct, np.uint32); co = np.ascontiguousarray(c_off, np.uint64)
        st = np.ascontiguousarray(strand, np.uint8); mo = np.ascontiguousarray(m_off, np.uint64)
        q = np.array(q, np.uint32); t = np.array(t, np.uint32)
        n = len(co) - 1
        keep = np.zeros(max(len(q), 1), np.uint8); removed = np.zeros(max(n, 1), np.int32)
        p = lambda x: _ptr(x) if x.size else None
        self._check(self.lib.lra_b200_trim_splitchains_batch(self.h, p(cq), p(ct), _ptr(co), p(st), n, p(q), p(t), _ptr(mo), _ptr(keep), _ptr(removed)))
        return q, t, keep[:len(q)], removed[:n]

    # ---- a22
    def mapq_batch(self, ag, bypass, read_type, global_k):
        """SetFromSegAlignment + AlignmentsOrder::Update + SimpleMapQV for every read.  ag: dict(grp_off, seg_off, upd_off, update_at, value, n0, n1, nm,
        nmm, ndel, nins, strand, flag, typeofaln, issec, supp).  Returns dict(flag, typeofaln, issec, supp, mapq, g_issec, g_value, g_n0, g_n1, g_nm, order)."""
        i32 = lambda k: np.ascontiguousarray(ag[k], np.int32)
        a = {k: i32(k) for k in ["grp_off", "seg_off", "upd_off", "update_at", "n0", "n1", "nm", "nmm", "ndel", "nins"]}
        a["value"] = np.ascontiguousarray(ag["value"], np.float32); a["strand"] = np.ascontiguousarray(ag["strand"], np.uint8)
        R = len(a["grp_off"]) - 1; G = int(a["grp_off"][-1]); S = int(a["seg_off"][G]) if G else 0
        o = dict(flag=np.array(ag["flag"], np.int32), typeofaln=np.array(ag["typeofaln"], np.int32), issec=np.array(ag["issec"], np.uint8), supp=np.array(ag["supp"], np.uint8),
                 mapq=np.zeros(max(S, 1), np.int32), g_issec=np.zeros(max(G, 1), np.uint8), g_value=np.zeros(max(G, 1), np.float32), g_n0=np.zeros(max(G, 1), np.int32),
                 g_n1=np.zeros(max(G, 1), np.int32), g_nm=np.zeros((max(G, 1), 4), np.int32), order=np.zeros(max(G, 1), np.int32))
        p = lambda x: _ptr(x) if x.size else None
        g = _AlignmentGroups(R, p(a["grp_off"]), p(a["seg_off"]), p(a["upd_off"]), p(a["update_at"]), p(a["value"]), p(a["n0"]), p(a["n1"]), p(a["nm"]), p(a["nmm"]),
                             p(a["ndel"]), p(a["nins"]), p(a["strand"]), bypass, read_type, global_k)
        r = _MapqResult(p(o["flag"]), p(o["typeofaln"]), p(o["issec"]), p(o["supp"]), _ptr(o["mapq"]), _ptr(o["g_issec"]), _ptr(o["g_value"]), _ptr(o["g_n0"]),
                        _ptr(o["g_n1"]), _ptr(o["g_nm"]), _ptr(o["order"]))
        self._check(self.lib.lra_b200_mapq_batch(self.h, C.byref(g), C.byref(r)))
        o["mapq"] = o["mapq"][:S]
        for k in ("g_issec", "g_value", "g_n0", "g_n1", "g_nm", "order"):
            o[k] = o[k][:G]
        return o

    # ---- a24
    def global_chain_batch(self, frag, frag_off, score):
        """GlobalChain for every problem (frag[n,4] int32, frag_off[n_prob+1], score[n]).  Returns dict(score, prev, chain, chain_len)."""
        f = np.ascontiguousarray(frag, np.int32).reshape(-1); fo = np.ascontiguousarray(frag_off, np.uint64)
        n = len(f) // 4
        o = dict(score=np.array(score, np.int32), prev=np.zeros(max(n, 1), np.int32), chain=np.zeros(max(n, 1), np.int32), chain_len=np.zeros(max(len(fo) - 1, 1), np.int32))
        if n == 0:
            o["score"] = np.zeros(1, np.int32)
        self._check(self.lib.lra_b200_global_chain_batch(self.h, _ptr(f) if n else None, _ptr(fo), len(fo) - 1, _ptr(o["score"]), _ptr(o["prev"]), _ptr(o["chain"]),
                                                         _ptr(o["chain_len"])))
        o["score"] = o["score"][:n]; o["prev"] = o["prev"][:n]; o["chain_len"] = o["chain_len"][:len(fo) - 1]
        return o

    # ---- a10
    def sdp_batch(self, pb, pwl, alnthres, num_aln, max_aln=2):
        """The SparseDP family for a batch of problems (tests/sdpgen.pack layout).  pwl = (stops, slope, inter, ceil1, ceil2) from init_pwl."""
        n = len(pb["mode"]); nf = int(pb["frag_off"][-1]) if n else 0
        o = dict(n_chains=np.zeros(max(n, 1), np.int32), chain_len=np.zeros(max(n * max_aln, 1), np.int32), chain_val=np.zeros(max(n * max_aln, 1), np.float32),
                 bounds=np.zeros(max(4 * n * max_aln, 1), np.uint32), chain=np.zeros(max(1, nf * max_aln), np.uint32), link=np.zeros(max(1, nf * max_aln), np.uint8),
                 cl_of_frag=np.zeros(max(1, nf), np.int32))
        keep = [np.ascontiguousarray(pwl[0], np.int64), np.ascontiguousarray(pwl[1], np.float32), np.ascontiguousarray(pwl[2], np.float32)]
        pr = _SdpProblems(n, max_aln, _ptr(pb["mode"]), _ptr(pb["frag_off"]), _ptr(pb["q"]), _ptr(pb["t"]), _ptr(pb["len"]), _ptr(pb["cl_off_off"]), _ptr(pb["cl_off"]),
                          _ptr(pb["cl_strand"]), _ptr(pb["only_cl"]), _ptr(pb["rate"]), _ptr(pb["irate"]), _ptr(pb["read_len"]), alnthres, num_aln,
                          _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), int(pwl[3]), int(pwl[4]),
                          *((_ptr(pb["qe"]), _ptr(pb["te"]), _ptr(pb["fstrand"]), _ptr(pb["fval"]), _ptr(pb["fn0"]), int(pb["globalK"])) if "qe" in pb else (None, None, None, None, None, 0)))
        o["n0"] = np.zeros(max(n * max_aln, 1), np.int32)
        rs = _SdpResult(_ptr(o["n_chains"]), _ptr(o["chain_len"]), _ptr(o["chain_val"]), _ptr(o["bounds"]), _ptr(o["chain"]), _ptr(o["link"]), _ptr(o["cl_of_frag"]), 0, _ptr(o["n0"]))
        self._check(self.lib.lra_b200_sdp_batch(self.h, C.byref(pr), C.byref(rs)))
        o["peak"] = int(rs.arena_peak); o["err"] = 0
        return o

    # ---- a12
    def lindex_build(self, seq, seq_start, seq_len, k=10, w=5, window=2048, max_freq=15, reuse=None):
        """LocalIndex::IndexSeq for every sequence [seq_start[s], +seq_len[s]) of a packed arena.  Returns a LocalIndexImage (`reuse`: an
        image of an earlier call, rebuilt in place)."""
        ss = np.ascontiguousarray(seq_start, np.uint64); sl = np.ascontiguousarray(seq_len, np.uint32)
        h = C.c_void_p(reuse.handle.value) if reuse is not None else C.c_void_p()
        if reuse is not None:
            reuse.handle = None
        self._check(self.lib.lra_b200_lindex_build(self.h, seq.handle, _ptr(ss), _ptr(sl), len(ss), k, w, window, max_freq, C.byref(h)))
        return LocalIndexImage(self, h)

    def gindex_build(self, seq, contig_start, contig_len, k=17, w=10, max_freq=150, win_size=15, per_window=1):
        """`lra index` / `lra global` on the device (StoreIndex, MMIndex.h:286-399).  Returns (t uint64[n], pos uint32[n]): the records of <ref>.mms."""
        cs = np.ascontiguousarray(contig_start, np.uint64); cl = np.ascontiguousarray(contig_len, np.uint32)
        h = C.c_void_p()
        self._check(self.lib.lra_b200_gindex_build(self.h, seq.handle, _ptr(cs), _ptr(cl), len(cs), k, w, max_freq, win_size, per_window, C.byref(h)))
        try:
            n = int(self.lib.lra_b200_index_size(h))
            t = np.zeros(n, np.uint64); pos = np.zeros(n, np.uint32)
            self._check(self.lib.lra_b200_index_download(self.h, h, _ptr(t) if n else None, _ptr(pos) if n else None))
        finally:
            self.lib.lra_b200_index_free(self.h, h)
        return t, pos

    def lindex_upload(self, seq_start, seq_len, window, win_off, bnd, mins):
        ss = np.ascontiguousarray(seq_start, np.uint64); sl = np.ascontiguousarray(seq_len, np.uint32)
        wo = np.ascontiguousarray(win_off, np.uint64); bd = np.ascontiguousarray(bnd, np.uint64); mn = np.ascontiguousarray(mins, np.uint32)
        h = C.c_void_p()
        self._check(self.lib.lra_b200_lindex_upload(self.h, _ptr(ss), _ptr(sl), len(ss), window, _ptr(wo), _ptr(bd), _ptr(mn) if len(mn) else None,
                                                    len(wo) - 1, C.byref(h)))
        return LocalIndexImage(self, h)

    # ---- a13
    def refine_clusters_batch(self, genome_li, reads_fwd, reads_rc, cl, anchor_cap=None):
        """REFINEclusters over a batch (cl: dict(m_q, m_t, m_off, box[n,4], strand, read_id, hdr_pos, global_k, small_k, window,
        local_max_freq)).  Returns dict(status, chrom, diag, r_off, r_q, r_t, r_tup, rbox, eff, m_q_out, m_t_out, box_out, n_anchors, ...)."""
        n = len(cl["strand"])
        a = dict(m_q=np.ascontiguousarray(cl["m_q"], np.uint32), m_t=np.ascontiguousarray(cl["m_t"], np.uint32), m_off=np.ascontiguousarray(cl["m_off"], np.uint64),
                 box=np.ascontiguousarray(cl["box"], np.uint32).reshape(-1), strand=np.ascontiguousarray(cl["strand"], np.uint8),
                 read_id=np.ascontiguousarray(cl["read_id"], np.uint32), hdr=np.ascontiguousarray(cl["hdr_pos"], np.uint64))
        M = int(a["m_off"][n]) if n else 0
        cap = anchor_cap if anchor_cap is not None else 4 * M + 4096
        for _ in range(2):
            o = dict(status=np.zeros(n, np.int32), chrom=np.zeros(n, np.int32), diag=np.zeros(2 * n, np.int64), r_off=np.zeros(n + 1, np.uint64),
                     r_q=np.zeros(cap, np.uint32), r_t=np.zeros(cap, np.uint32), r_tup=np.zeros(cap, np.uint32), rbox=np.zeros(4 * n, np.uint32),
                     eff=np.zeros(n, np.float32), m_q_out=np.zeros(M + 1, np.uint32), m_t_out=np.zeros(M + 1, np.uint32), box_out=np.zeros(4 * n, np.uint32))
            c = _Clusters(n, _ptr(a["m_q"]) if M else None, _ptr(a["m_t"]) if M else None, _ptr(a["m_off"]), _ptr(a["box"]), _ptr(a["strand"]),
                          _ptr(a["read_id"]), _ptr(a["hdr"]), len(a["hdr"]), cl["global_k"], cl["small_k"], cl["window"], cl["local_max_freq"])
            r = _Refined(_ptr(o["status"]), _ptr(o["chrom"]), _ptr(o["diag"]), _ptr(o["r_off"]), _ptr(o["r_q"]), _ptr(o["r_t"]), _ptr(o["r_tup"]), cap, 0,
                         _ptr(o["rbox"]), _ptr(o["eff"]), _ptr(o["m_q_out"]), _ptr(o["m_t_out"]), _ptr(o["box_out"]), 0, 0)
            rc = self.lib.lra_b200_refine_clusters_batch(self.h, genome_li.handle, reads_fwd.handle, reads_rc.handle, C.byref(c), C.byref(r))
            o["n_anchors"], o["n_units"], o["n_tasks"] = int(r.n_anchors), int(r.n_units), int(r.n_tasks)
            if rc == EOVERFLOW and anchor_cap is None:
                cap = int(r.n_anchors) + 16
                continue
            self._check(rc)
            return o
        self._check(rc)

    def refine_clusters_batch_device(self, genome_li, reads_fwd, reads_rc, p, n, n_anchors_in, consts, out, anchor_cap):
        """Device-pointer variant.  p: dict of device pointers m_q, m_t, m_off, box, strand, read_id, hdr_pos (+ n_hdr); consts: (global_k,
        small_k, window, local_max_freq); out: dict of device pointers status, chrom, diag, r_off, r_q, r_t, r_tup, rbox, eff.
        Returns dict(n_anchors, n_units, n_tasks)."""
        c = _Clusters(n, p["m_q"], p["m_t"], p["m_off"], p["box"], p["strand"], p["read_id"], p["hdr_pos"], p["n_hdr"], consts[0], consts[1], consts[2], consts[3])
        r = _Refined(out["status"], out["chrom"], out["diag"], out["r_off"], out["r_q"], out["r_t"], out["r_tup"], anchor_cap, 0, out["rbox"], out["eff"],
                     None, None, None, 0, 0)
        self._check(self.lib.lra_b200_refine_clusters_batch_device(self.h, genome_li.handle, reads_fwd.handle, reads_rc.handle, C.byref(c), n_anchors_in, C.byref(r)))
        return dict(n_anchors=int(r.n_anchors), n_units=int(r.n_units), n_tasks=int(r.n_tasks))

    def refine_splitchains_batch(self, genome_li, reads_fwd, reads_rc, sc, anchor_cap=None, out=None):
        """Refine_splitchain over a batch (sc: dict(m_q, m_t, m_len, m_strand, m_off, box[n,4], strand, chrom, read_id, hdr_pos, global_k, small_k,
        window, local_max_freq, limitrefine)).  Returns dict(status, chrom, diag, r_off, r_q, r_t, r_tup, rbox, eff, n_anchors, n_units, n_tasks)."""
        n = len(sc["strand"])
        a = dict(m_q=np.ascontiguousarray(sc["m_q"], np.uint32), m_t=np.ascontiguousarray(sc["m_t"], np.uint32), m_len=np.ascontiguousarray(sc["m_len"], np.uint32),
                 m_strand=np.ascontiguousarray(sc["m_strand"], np.uint8), m_off=np.ascontiguousarray(sc["m_off"], np.uint64),
                 box=np.ascontiguousarray(sc["box"], np.uint32).reshape(-1), strand=np.ascontiguousarray(sc["strand"], np.uint8),
                 chrom=np.ascontiguousarray(sc["chrom"], np.int32), read_id=np.ascontiguousarray(sc["read_id"], np.uint32),
                 hdr=np.ascontiguousarray(sc["hdr_pos"], np.uint64))
        M = int(a["m_off"][n]) if n else 0
        cap = anchor_cap if anchor_cap is not None else 4 * M + 4096
        for _ in range(2):
            # `out`: caller-owned (e.g. pinned) result arrays of the right sizes, re-used across batches
            o = out if out is not None else dict(status=np.zeros(n, np.int32), chrom=np.zeros(n, np.int32), diag=np.zeros(2 * n, np.int64), r_off=np.zeros(n + 1, np.uint64),
                                                 r_q=np.zeros(cap, np.uint32), r_t=np.zeros(cap, np.uint32), r_tup=np.zeros(cap, np.uint32), rbox=np.zeros(4 * n, np.uint32),
                                                 eff=np.zeros(n, np.float32))
            c = _SplitChains(n, _ptr(a["m_q"]) if M else None, _ptr(a["m_t"]) if M else None, _ptr(a["m_len"]) if M else None, _ptr(a["m_strand"]) if M else None,
                             _ptr(a["m_off"]), _ptr(a["box"]), _ptr(a["strand"]), _ptr(a["chrom"]), _ptr(a["read_id"]), _ptr(a["hdr"]), len(a["hdr"]),
                             sc["global_k"], sc["small_k"], sc["window"], sc["local_max_freq"], sc.get("limitrefine", 1))
            r = _Refined(_ptr(o["status"]), _ptr(o["chrom"]), _ptr(o["diag"]), _ptr(o["r_off"]), _ptr(o["r_q"]), _ptr(o["r_t"]), _ptr(o["r_tup"]), cap, 0,
                         _ptr(o["rbox"]), _ptr(o["eff"]), None, None, None, 0, 0)
            rc = self.lib.lra_b200_refine_splitchains_batch(self.h, genome_li.handle, reads_fwd.handle, reads_rc.handle, C.byref(c), C.byref(r))
            o["n_anchors"], o["n_units"], o["n_tasks"] = int(r.n_anchors), int(r.n_units), int(r.n_tasks)
            if rc == EOVERFLOW and anchor_cap is None:
                cap = int(r.n_anchors) + 16
                continue
            self._check(rc)
            return o
        self._check(rc)

    def refine_splitchains_batch_device(self, genome_li, reads_fwd, reads_rc, p, n, n_anchors_in, consts, out, anchor_cap):
        """Device-pointer variant.  p: device pointers m_q, m_t, m_len, m_strand, m_off, box, strand, chrom, read_id, hdr_pos (+ n_hdr); consts:
        (global_k, small_k, window, local_max_freq, limitrefine); out: device pointers status, chrom, diag, r_off, r_q, r_t, r_tup, rbox, eff."""
        c = _SplitChains(n, p["m_q"], p["m_t"], p["m_len"], p["m_strand"], p["m_off"], p["box"], p["strand"], p["chrom"], p["read_id"], p["hdr_pos"], p["n_hdr"],
                         consts[0], consts[1], consts[2], consts[3], consts[4])
        r = _Refined(out["status"], out["chrom"], out["diag"], out["r_off"], out["r_q"], out["r_t"], out["r_tup"], anchor_cap, 0, out["rbox"], out["eff"],
                     None, None, None, 0, 0)
        self._check(self.lib.lra_b200_refine_splitchains_batch_device(self.h, genome_li.handle, reads_fwd.handle, reads_rc.handle, C.byref(c), n_anchors_in, C.byref(r)))
        return dict(n_anchors=int(r.n_anchors), n_units=int(r.n_units), n_tasks=int(r.n_tasks))

    def calc_stats_batch_device(self, q, t, ptrs, n_blocks_in, S, log_lut, d_stats, d_value, d_cigar_off, d_cigar, cigar_cap):
        """Device-pointer variant.  ptrs = [blocks, blk_off, blk_cnt, q_base, t_base, read_len] (blk_off uint64, blk_cnt int32: the
        block_off / n_blocks outputs of indel_refine_batch_device fit).  Returns the number of CIGAR ops."""
        lut = np.ascontiguousarray(log_lut, np.float32)
        sg = _IrSegments(ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], None, n_blocks_in, S, 0, 0, 0, 0, 0)
        res = _StatsResult(d_stats, d_value, d_cigar_off, d_cigar, cigar_cap, 0)
        self._check(self.lib.lra_b200_calc_stats_batch_device(self.h, q.handle, t.handle, C.byref(sg), _ptr(lut), C.byref(res)))
        return int(res.n_cigar_total)

    # ---- a21
    def calc_stats_batch(self, q, t, sb, log_lut, cigar_cap=None, out=None):
        """Alignment::CalculateStatistics over segments (sb as for indel_refine_batch: blocks_in, blk_off, blk_cnt, q_base, t_base,
        read_len).  Returns dict(stats[S,16], value[S], cigar_off[S+1], cigar)."""
        S = len(sb["blk_cnt"])
        bi = np.ascontiguousarray(sb["blocks_in"], np.uint32)
        a = dict(blk_off=np.ascontiguousarray(sb["blk_off"], np.uint64), blk_cnt=np.ascontiguousarray(sb["blk_cnt"], np.int32),
                 q_base=np.ascontiguousarray(sb["q_base"], np.uint32), t_base=np.ascontiguousarray(sb["t_base"], np.uint32),
                 read_len=np.ascontiguousarray(sb["read_len"], np.int32))
        cl = np.zeros(S, np.int32)
        lut = np.ascontiguousarray(log_lut, np.float32)
        assert len(lut) == 2001
        cap = cigar_cap if cigar_cap is not None else 4 * (bi.size // 3) + 16 * S + 16
        o = out if out is not None else dict(stats=np.zeros((S, 16), np.int32), value=np.zeros(S, np.float32), cigar_off=np.zeros(S + 1, np.uint64), cigar=np.zeros(cap, np.uint32))
        sg = _IrSegments(_ptr(bi), _ptr(a["blk_off"]), _ptr(a["blk_cnt"]), _ptr(a["q_base"]), _ptr(a["t_base"]), _ptr(a["read_len"]), _ptr(cl),
                         bi.size // 3, S, 0, 0, 0, 0, 0)
        res = _StatsResult(_ptr(o["stats"]), _ptr(o["value"]), _ptr(o["cigar_off"]), _ptr(o["cigar"]), cap, 0)
        rc = self.lib.lra_b200_calc_stats_batch(self.h, q.handle, t.handle, C.byref(sg), _ptr(lut), C.byref(res))
        o["n_cigar_total"] = int(res.n_cigar_total)
        self._check(rc)
        return o

    # ---- a19
    def indel_dp_batch(self, q, t, g, block_cap=None, out=None):
        """g: dict of host arrays (q_base,t_base,q_start,t_start,t_len,q_seq_len,t_seq_len,band_off,band) + match/mismatch/indel.
        Returns dict(n_blocks, block_off, blocks[n,3], n_blocks_total, cells)."""
        n = len(g["t_len"])
        arr = {k: np.ascontiguousarray(g[k], np.uint32 if k in ("q_base", "t_base", "band_off") else np.int32)
               for k in ("q_base", "t_base", "q_start", "t_start", "t_len", "q_seq_len", "t_seq_len", "band_off", "band")}
        if block_cap is None:
            block_cap = int(arr["q_seq_len"].sum() + arr["t_seq_len"].sum()) + 8
        if out is None:
            out = dict(n_blocks=np.zeros(n, np.int32), block_off=np.zeros(n, np.uint64), blocks=np.zeros((max(1, block_cap), 3), np.uint32))
        gs = _IrGroups(*[_ptr(arr[k]) for k in ("q_base", "t_base", "q_start", "t_start", "t_len", "q_seq_len", "t_seq_len", "band_off", "band")],
                       len(arr["band"]), n, g["match"], g["mismatch"], g["indel"])
        res = _IrResult(_ptr(out["n_blocks"]), _ptr(out["block_off"]), _ptr(out["blocks"]), block_cap, 0, 0)
        rc = self.lib.lra_b200_indel_dp_batch(self.h, q.handle, t.handle, C.byref(gs), C.byref(res))
        out["n_blocks_total"] = int(res.n_blocks_total); out["cells"] = int(res.cells)
        self._check(rc)
        return out

    def indel_refine_batch(self, q, t, sb, block_cap=None, out=None):
        """Whole IndelRefineAlignment over segments.  sb: dict(blocks_in[T,3], blk_off, blk_cnt, q_base, t_base, read_len,
        contig_len, k, match, mismatch, indel, end_align).  Returns dict(n_blocks, block_off, blocks, ...)."""
        S = len(sb["blk_cnt"])
        bi = np.ascontiguousarray(sb["blocks_in"], np.uint32)
        a = dict(blk_off=np.ascontiguousarray(sb["blk_off"], np.uint64), blk_cnt=np.ascontiguousarray(sb["blk_cnt"], np.int32),
                 q_base=np.ascontiguousarray(sb["q_base"], np.uint32), t_base=np.ascontiguousarray(sb["t_base"], np.uint32),
                 read_len=np.ascontiguousarray(sb["read_len"], np.int32), contig_len=np.ascontiguousarray(sb["contig_len"], np.int32))
        if block_cap is None:
            block_cap = int(a["read_len"].sum()) + 16
        if out is None:
            out = dict(n_blocks=np.zeros(S, np.int32), block_off=np.zeros(S, np.uint64), blocks=np.zeros((max(1, block_cap), 3), np.uint32))
        sg = _IrSegments(_ptr(bi), _ptr(a["blk_off"]), _ptr(a["blk_cnt"]), _ptr(a["q_base"]), _ptr(a["t_base"]), _ptr(a["read_len"]),
                         _ptr(a["contig_len"]), bi.size // 3, S, sb["k"], sb["match"], sb["mismatch"], sb["indel"], sb["end_align"])
        res = _IrSegResult(_ptr(out["n_blocks"]), _ptr(out["block_off"]), _ptr(out["blocks"]), block_cap, 0, 0, 0, 0)
        rc = self.lib.lra_b200_indel_refine_batch(self.h, q.handle, t.handle, C.byref(sg), C.byref(res))
        out.update(n_blocks_total=int(res.n_blocks_total), cells=int(res.cells), n_dp_groups=int(res.n_dp_groups), n_aog_jobs=int(res.n_aog_jobs))
        self._check(rc)
        return out

    def indel_refine_batch_device(self, q, t, ptrs, n_blocks_in, S, k, match, mismatch, indel, end_align, d_n_blocks, d_block_off,
                                  d_blocks, block_cap):
        """ptrs: device pointers blocks_in, blk_off, blk_cnt, q_base, t_base, read_len, contig_len."""
        sg = _IrSegments(*ptrs, n_blocks_in, S, k, match, mismatch, indel, end_align)
        res = _IrSegResult(d_n_blocks, d_block_off, d_blocks, block_cap, 0, 0, 0, 0)
        self._check(self.lib.lra_b200_indel_refine_batch_device(self.h, q.handle, t.handle, C.byref(sg), C.byref(res)))
        return dict(n_blocks_total=int(res.n_blocks_total), cells=int(res.cells), n_dp_groups=int(res.n_dp_groups), n_aog_jobs=int(res.n_aog_jobs))

    def indel_dp_batch_device(self, q, t, ptrs, n, band_len, match, mismatch, indel, d_n_blocks, d_block_off, d_blocks, block_cap):
        """ptrs: device pointers in the order q_base,t_base,q_start,t_start,t_len,q_seq_len,t_seq_len,band_off,band."""
        gs = _IrGroups(*ptrs, band_len, n, match, mismatch, indel)
        res = _IrResult(d_n_blocks, d_block_off, d_blocks, block_cap, 0, 0)
        self._check(self.lib.lra_b200_indel_dp_batch_device(self.h, q.handle, t.handle, C.byref(gs), C.byref(res)))
        return int(res.n_blocks_total), int(res.cells)


# ---------------------------------------------------------------------------------------------------------------------
# Host-side mirror of the reference interface (same names / argument meaning), used by the parity tests.

_default_ctx = None


def CreateLookUpTable():
    """The reference's LogLookUpTable (LogLookUpTable.h:9-15): logf(i) for i = 1, 6, ..., 10001, built with the HOST libm so that the NV
    tag is bit-identical (never a device logf for anything that reaches the output)."""
    libm = C.CDLL("libm.so.6")
    libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
    return np.array([libm.logf(float(i)) for i in range(1, 10002, 5)], np.float32)


def write_gli(path, k, w, window, seq_offsets, tuple_boundaries, minimizers):
    """<ref>.gli as LocalIndex::Write lays it out (MMIndex.h:138-151): int k, w, localIndexWindow, nRegions; uint64 seqOffsets[nRegions];
    uint64 tupleBoundaries[nRegions]; uint64 nMin; LocalTuple minimizers[nMin] (uint32: tuple | pos << 20).  seq_offsets / tuple_boundaries
    include the leading 0 (what LocalIndexImage.download() returns for an arena that starts at 0)."""
    so = np.ascontiguousarray(seq_offsets, np.uint64); tb = np.ascontiguousarray(tuple_boundaries, np.uint64); mn = np.ascontiguousarray(minimizers, np.uint32)
    assert len(so) == len(tb)
    with open(path, "wb") as f:
        f.write(np.array([k, w, window, len(so)], np.int32).tobytes())
        f.write(so.tobytes()); f.write(tb.tobytes())
        f.write(np.array([len(mn)], np.uint64).tobytes()); f.write(mn.tobytes())


INDEX_PRESET = {   # lra.cpp:884-911 (global) and the LocalIndex the same command writes (k 10, w 5, window 2048, maxFreq 15)
    "ont": dict(k=17, w=10, max_freq=150, win_size=15, per_window=1), "ccs": dict(k=17, w=10, max_freq=150, win_size=15, per_window=1),
    "clr": dict(k=15, w=10, max_freq=250, win_size=12, per_window=1), "contig": dict(k=19, w=10, max_freq=30, win_size=20, per_window=1)}


def write_mms(path, k, names, hdr, t, pos):
    """<ref>.mms as WriteIndex lays it out (MMIndex.h:414-422): int64 n, int k, Header (int nContigs; per contig int len + name; uint64 pos[n+1]),
    GenomeTuple[n] (uint64 t, uint32 pos, 4 bytes of padding)."""
    t = np.ascontiguousarray(t, np.uint64); pos = np.ascontiguousarray(pos, np.uint32)
    with open(path, "wb") as f:
        f.write(np.array([len(t)], np.int64).tobytes()); f.write(np.array([k], np.int32).tobytes())
        f.write(np.array([len(names)], np.int32).tobytes())
        for n in names:
            b = n.encode(); f.write(np.array([len(b)], np.int32).tobytes()); f.write(b)
        f.write(np.ascontiguousarray(hdr, np.uint64).tobytes())
        CH = 1 << 24
        for a in range(0, len(t), CH):
            rec = np.zeros(min(CH, len(t) - a), np.dtype([("t", "<u8"), ("pos", "<u4"), ("pad", "<u4")]))
            rec["t"] = t[a:a + CH]; rec["pos"] = pos[a:a + CH]
            f.write(rec.tobytes())


def build_index_files(fasta_path, contigs, preset, device=0, ctx=None):
    """`lra index -MODE ref.fa` on the GPU: writes <fasta>.mms and <fasta>.gli in the reference's formats (the reference loads them).
    contigs: list of (name, uint8 ASCII array) as in the FASTA.  Returns dict(n_mms, n_gli)."""
    own = ctx is None
    ctx = ctx or Context(device)
    pr = INDEX_PRESET[preset]
    lens = np.array([len(s) for _, s in contigs], np.uint32)
    start = np.zeros(len(contigs), np.uint64); start[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    hdr = np.concatenate([start, [np.uint64(int(start[-1]) + int(lens[-1]))]]).astype(np.uint64)
    arena = ctx.seq_upload(np.concatenate([s for _, s in contigs]))
    try:
        t, pos = ctx.gindex_build(arena, start, lens, **pr)
        write_mms(fasta_path + ".mms", pr["k"], [n for n, _ in contigs], hdr, t, pos)
        img = ctx.lindex_build(arena, start, lens, k=10, w=5, window=2048, max_freq=15)
        wo, bd, mn = img.download()
        write_gli(fasta_path + ".gli", 10, 5, 2048, wo, bd, mn)
        img.free()
    finally:
        arena.free()
        if own:
            ctx.close()
    return dict(n_mms=len(t), n_gli=len(mn))


def read_gli(path):
    """Inverse of write_gli (LocalIndex::Read, MMIndex.h:154-173).  Returns dict(k, w, window, seq_offsets, tuple_boundaries, minimizers)."""
    d = open(path, "rb").read()
    k, w, window, n = (int(x) for x in np.frombuffer(d, np.int32, 4, 0))
    so = np.frombuffer(d, np.uint64, n, 16).copy(); tb = np.frombuffer(d, np.uint64, n, 16 + 8 * n).copy()
    nm = int(np.frombuffer(d, np.uint64, 1, 16 + 16 * n)[0])
    return dict(k=k, w=w, window=window, seq_offsets=so, tuple_boundaries=tb, minimizers=np.frombuffer(d, np.uint32, nm, 24 + 16 * n).copy())


def _ctx():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def AffineOneGapAlign(qSeq, qLen, tSeq, tLen, m, mm, indel, k, ctx=None):
    """Mirror of `int AffineOneGapAlign(string &qSeq, int qLen, string &tSeq, int tLen, int m, int mm, int indel, int k,
    Alignment &aln, AffineAlignBuffers &b)` (reference AffineOneGapAlign.h:157).  Returns (score, blocks) where blocks
    is what the reference appends to aln.blocks, as an [n,3] array of (qPos, tPos, length)."""
    ctx = ctx or _ctx()
    q = ctx.seq_upload(bytes(qSeq[:qLen])); t = ctx.seq_upload(bytes(tSeq[:tLen]))
    try:
        r = ctx.aog_batch(q, t, [0], [0], [qLen], [tLen], [k], m, mm, indel)
    finally:
        q.free(); t.free()
    nb = int(r["n_blocks"][0]); off = int(r["block_off"][0])
    return int(r["score"][0]), r["blocks"][off:off + nb].copy()


def IndelRefineAlignment(read, tSeq, contigLen, blocks, refineBand, localMatch, localMismatch, localIndel, endAlign=False, tWinOff=0,
                         ctx=None):
    """Mirror of `void IndelRefineAlignment(Read &read, Genome &genome, Alignment &alignment, const Options &opts,
    IndelRefineBuffers &buffers, bool endAlign)` (reference IndelRefine.h:53) for one segment: `read` is alignment.read (the
    strand the blocks refer to), `tSeq` the contig (or a window of it starting at contig offset tWinOff), `blocks` the
    segment's alignment.blocks.  Returns the refined blocks."""
    ctx = ctx or _ctx()
    q = ctx.seq_upload(bytes(read)); t = ctx.seq_upload(bytes(tSeq))
    try:
        b = np.ascontiguousarray(blocks, np.uint32).reshape(-1, 3)
        sb = dict(blocks_in=b, blk_off=np.zeros(1, np.uint64), blk_cnt=np.array([len(b)], np.int32), q_base=np.zeros(1, np.uint32),
                  t_base=np.array([(-tWinOff) & 0xFFFFFFFF], np.uint32), read_len=np.array([len(read)], np.int32),
                  contig_len=np.array([contigLen], np.int32), k=refineBand, match=localMatch, mismatch=localMismatch,
                  indel=localIndel, end_align=1 if endAlign else 0)
        r = ctx.indel_refine_batch(q, t, sb)
    finally:
        q.free(); t.free()
    return r["blocks"][int(r["block_off"][0]):int(r["block_off"][0]) + int(r["n_blocks"][0])].copy()


def AffineOneGapAlignBatch(q_arena, t_arena, q_off, t_off, q_len, t_len, k, m, mm, indel, ctx=None):
    """Batched form over SoA job arrays (ASCII arenas as bytes / uint8 arrays)."""
    ctx = ctx or _ctx()
    q = ctx.seq_upload(q_arena); t = ctx.seq_upload(t_arena)
    try:
        return ctx.aog_batch(q, t, q_off, t_off, q_len, t_len, k, m, mm, indel)
    finally:
        q.free(); t.free()
