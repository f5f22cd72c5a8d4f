"""ctypes binding of the C ABI declared in ``include/ppsurf_b200.h``.  There is no fallback: if the shared library is
missing or fails to load, importing this module raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libppsurf_b200.so')

c_f32p = ctypes.c_void_p
c_i32p = ctypes.c_void_p
c_voidp = ctypes.c_void_p
i64 = ctypes.c_int64
i32 = ctypes.c_int
size_t = ctypes.c_size_t


class PpsError(RuntimeError):
    pass


class DecoderWeights(ctypes.Structure):
    """mirror of ``pps_decoder_weights``"""
    _fields_ = [
        ('latent', ctypes.c_int32), ('heads', ctypes.c_int32), ('k', ctypes.c_int32), ('num_pts_local', ctypes.c_int32),
        ('w1_lat', c_f32p), ('w1_xyz', c_f32p), ('b1', c_f32p), ('w2', c_f32p), ('b2', c_f32p), ('w3', c_f32p),
        ('b3', c_f32p), ('wq', c_f32p), ('bq', c_f32p), ('wv8', c_f32p), ('bv8', c_f32p),
        ('pn0a_w', c_f32p), ('pn0a_b', c_f32p), ('pn0b_w', c_f32p), ('pn0b_b', c_f32p),
        ('stn1_w', c_f32p), ('stn1_b', c_f32p), ('stn2_w', c_f32p), ('stn2_b', c_f32p), ('stn3_w', c_f32p),
        ('stn3_b', c_f32p), ('stnf1_w', c_f32p), ('stnf1_b', c_f32p), ('stnf2_w', c_f32p), ('stnf2_b', c_f32p),
        ('stnf3_w', c_f32p), ('stnf3_b', c_f32p), ('pn1_w', c_f32p), ('pn1_b', c_f32p), ('pn2_w', c_f32p),
        ('pn2_b', c_f32p), ('pnq_w', c_f32p), ('pnq_b', ctypes.c_float), ('stn_size', ctypes.c_int32),
        ('pnv_w', c_f32p), ('pnv_b', c_f32p),
        ('m0_w', c_f32p), ('m0_b', c_f32p), ('m1_w', c_f32p), ('m1_b', c_f32p), ('m2_w', c_f32p), ('m2_b', c_f32p),
        ('tc_wpack', c_voidp), ('tc_pn_stn', c_voidp), ('tc_pn_feat', c_voidp), ('tc_stn_fc', c_voidp), ('tc_mlp', c_voidp),
        ('tc_bias_feat', c_f32p),
    ]


class FKAConvWeights(ctypes.Structure):
    """mirror of ``pps_fkaconv_weights``"""
    _fields_ = [
        ('cin', ctypes.c_int32), ('cout', ctypes.c_int32), ('act', ctypes.c_int32),
        ('alpha', ctypes.c_float), ('beta', ctypes.c_float), ('norm_radius', ctypes.c_float),
        ('fc1', c_f32p), ('fc2', c_f32p), ('fc3', c_f32p), ('in1_w', c_f32p), ('in1_b', c_f32p), ('in2_w', c_f32p),
        ('in2_b', c_f32p), ('cv_w', c_f32p), ('out_bias', c_f32p), ('out_relu', ctypes.c_int32), ('tc_pack', c_voidp), ('mlp_host', c_voidp), ('tc_out_scale', ctypes.c_float),
    ]


class EncoderIdsOut(ctypes.Structure):
    """mirror of ``pps_encoder_ids_out``"""
    _fields_ = [('support', c_f32p * 4), ('ids16', c_i32p * 9), ('ids1', c_i32p * 4)]


# name -> (restype, argtypes); every symbol include/ppsurf_b200.h declares
SIGNATURES = {
    'pps_last_error': (ctypes.c_char_p, []),
    'pps_version': (i32, []),
    'pps_source_hash': (ctypes.c_ulonglong, []),
    'pps_compiled_arch': (i32, []),
    'pps_check_device': (i32, []),
    'pps_launch_count': (ctypes.c_ulonglong, []),
    'pps_profile_enable': (None, [i32]),
    'pps_profile_read': (i32, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
    'pps_knn_index_bytes': (size_t, [i64]),
    'pps_knn_build': (i32, [c_f32p, i64, c_voidp, size_t, c_voidp]),
    'pps_debug_knn_run': (i32, [i32]),
    'pps_debug_knn_cells': (i32, [i32]),
    'pps_debug_knn_scan_child': (i32, [i32]),
    'pps_debug_knn_scan_cap': (i32, [i32]),
    'pps_knn_query': (i32, [c_voidp, i64, c_f32p, i64, i32, c_i32p, c_f32p, c_voidp]),
    'pps_patch_normalize': (i32, [c_f32p, c_f32p, c_i32p, c_f32p, i64, i32, i32, c_f32p, c_voidp]),
    'pps_linear': (i32, [c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_f32p, i64, i32, i32, i32, i32, i32, c_voidp]),
    'pps_decoder_point_table': (i32, [ctypes.POINTER(DecoderWeights), c_f32p, c_f32p, i64, c_f32p, c_voidp]),
    'pps_decoder_tc_pack_bytes': (size_t, []),
    'pps_debug_tc_profile': (None, [c_voidp]),
    'pps_debug_tc_cluster': (None, [i32]),
    'pps_decoder_tc_terms': (i32, [i32]),
    'pps_debug_tc_max_clusters': (i32, []),
    'pps_decoder_tc_pn_stn_bytes': (size_t, []),
    'pps_decoder_tc_pn_feat_bytes': (size_t, []),
    'pps_decoder_tc_stn_fc_bytes': (size_t, []),
    'pps_decoder_tc_mlp_bytes': (size_t, []),
    'pps_decoder_workspace_bytes': (size_t, [ctypes.POINTER(DecoderWeights), i64]),
    'pps_decoder_decode': (i32, [ctypes.POINTER(DecoderWeights), c_voidp, c_f32p, c_f32p, i64, c_f32p, i64, i64, c_voidp,
                                 size_t, c_f32p, c_f32p, c_i32p, i32, c_voidp]),
    'pps_decoder_decode_host': (i32, [ctypes.POINTER(DecoderWeights), c_voidp, c_f32p, c_f32p, i64, c_voidp, i64, i64,
                                      c_voidp, size_t, c_voidp, size_t, c_voidp, i32, c_voidp, c_voidp]),
    'pps_decoder_projection': (i32, [ctypes.POINTER(DecoderWeights), c_f32p, c_f32p, c_f32p, c_i32p, i32, i64, c_voidp,
                                     size_t, c_f32p, i32, c_voidp]),
    'pps_decoder_pointnet': (i32, [ctypes.POINTER(DecoderWeights), c_f32p, i64, c_voidp, size_t, c_f32p, i32, c_voidp]),
    'pps_grid_queries': (i32, [i32, ctypes.c_float, ctypes.c_float, i64, i64, c_f32p, c_voidp]),
    'pps_region_workspace_bytes': (size_t, [i32]),
    'pps_region_init': (i32, [i32, c_f32p, c_voidp, c_voidp]),
    'pps_region_pending': (i32, [c_i32p, i64, i32, i32, c_f32p, c_voidp, size_t, c_i32p, c_voidp, c_voidp]),
    'pps_region_queries': (i32, [c_i32p, i64, i32, ctypes.c_float, ctypes.c_float, c_f32p, c_voidp]),
    'pps_region_scatter': (i32, [c_i32p, c_f32p, i64, c_f32p, c_voidp]),
    'pps_region_frontier': (i32, [c_i32p, i64, i32, i32, c_f32p, c_voidp, c_voidp, size_t, c_i32p, c_voidp, c_voidp]),
    'pps_region_finish': (i32, [c_f32p, i32, i32, ctypes.c_float, c_voidp]),
    'pps_mc_set_edges': (i32, [c_voidp, c_voidp, c_voidp]),
    'pps_mc_workspace_bytes': (size_t, [i32]),
    'pps_mc_count': (i32, [c_f32p, i32, ctypes.c_float, c_voidp, i32, c_voidp, size_t, c_voidp, c_voidp]),
    'pps_mc_emit': (i32, [c_f32p, i32, ctypes.c_float, c_voidp, i32, c_voidp, c_f32p, c_i32p, c_i32p, c_voidp]),
    'pps_refine_init': (i32, [c_f32p, i32, c_f32p, c_i32p, i64, ctypes.c_float, ctypes.c_float, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                              c_voidp, c_voidp]),
    'pps_refine_update': (i32, [c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_voidp]),
    'pps_sample_workspace_bytes': (size_t, [i64]),
    'pps_sample_quantized': (i32, [c_f32p, i64, i64, c_f32p, i32, ctypes.c_uint32, c_voidp, size_t, c_i32p, c_voidp]),
    'pps_encoder_ids_workspace_bytes': (size_t, [i64]),
    'pps_encoder_ids': (i32, [c_f32p, i64, i64, c_f32p, i32, ctypes.c_uint32, c_voidp, size_t, ctypes.POINTER(EncoderIdsOut), c_voidp]),
    'pps_fkaconv_workspace_bytes': (size_t, [i64, i64, i32]),
    'pps_fkaconv_workspace_bytes_for': (size_t, [ctypes.POINTER(FKAConvWeights), i32, i64, i64]),
    'pps_debug_fka_fused': (None, [i32]),
    'pps_fkaconv_forward': (i32, [ctypes.POINTER(FKAConvWeights), c_f32p, c_f32p, c_f32p, c_i32p, i32, i64, i64, i64, c_voidp,
                                  size_t, c_f32p, c_voidp]),
    'pps_gather_max': (i32, [c_f32p, c_i32p, i64, i64, i64, i32, i32, c_f32p, c_voidp]),
    'pps_global_max': (i32, [c_f32p, i64, i64, i32, c_f32p, c_voidp]),
    'pps_latent_accumulate': (i32, [c_f32p, c_i32p, i64, i32, c_f32p, c_f32p, c_voidp]),
    'pps_latent_accumulate_rows': (i32, [c_f32p, c_i32p, c_i32p, i64, i32, c_f32p, c_f32p, c_voidp]),
    'pps_latent_finalize': (i32, [c_f32p, c_f32p, i64, i32, c_voidp]),
    # ---- config 5: training primitives (csrc/train_gemm.cu, train_gemm_tc.cu, train_ops.cu)
    'pps_gemm': (i32, [c_f32p, i64, i64, i64, c_f32p, i64, i64, i64, c_f32p, i64, i64, i64, i64, i32, i64, c_f32p, i32, i32, c_voidp]),
    'pps_colsum': (i32, [c_f32p, i64, i32, i64, c_f32p, i32, c_voidp]),
    'pps_norm_workspace_bytes': (size_t, [i64, i32]),
    'pps_norm_fwd': (i32, [c_f32p, i64, i64, i32, c_f32p, c_f32p, ctypes.c_float, i32, c_f32p, c_f32p, c_f32p, c_voidp, size_t, c_voidp]),
    'pps_norm_bwd': (i32, [c_f32p, c_f32p, i64, i64, i32, c_f32p, c_f32p, c_f32p, c_f32p, ctypes.c_float, i32, c_f32p, c_f32p, c_f32p,
                           c_voidp, size_t, c_voidp]),
    'pps_bn_running_update': (i32, [c_f32p, c_f32p, i64, ctypes.c_float, i32, c_f32p, c_f32p, c_voidp]),
    'pps_act_fwd': (i32, [c_f32p, i64, i32, c_f32p, c_voidp]),
    'pps_act_bwd': (i32, [c_f32p, c_f32p, i64, i32, c_f32p, c_voidp]),
    'pps_dropout_fwd': (i32, [c_f32p, i64, ctypes.c_float, ctypes.c_uint32, c_voidp, c_f32p, c_voidp, c_voidp]),
    'pps_dropout_bwd': (i32, [c_f32p, c_voidp, i64, ctypes.c_float, c_f32p, c_voidp]),
    'pps_rowscale_fwd': (i32, [c_f32p, c_f32p, i64, i32, c_f32p, c_voidp]),
    'pps_rowscale_bwd': (i32, [c_f32p, c_f32p, c_f32p, i64, i32, c_f32p, c_f32p, c_voidp]),
    'pps_concat_bcast_fwd': (i32, [c_f32p, c_f32p, i64, i32, i32, c_f32p, c_voidp]),
    'pps_concat_bcast_bwd': (i32, [c_f32p, i64, i32, i32, c_f32p, c_f32p, c_voidp]),
    'pps_gather_rows': (i32, [c_f32p, c_i32p, i64, i32, c_f32p, c_voidp]),
    'pps_scatter_add_rows': (i32, [c_f32p, c_i32p, i64, i32, c_f32p, c_voidp]),
    'pps_seg_max_fwd': (i32, [c_f32p, c_f32p, i64, i32, i32, c_f32p, c_i32p, c_voidp]),
    'pps_seg_max_bwd': (i32, [c_f32p, c_i32p, c_f32p, c_f32p, i64, i32, i32, c_f32p, c_f32p, c_voidp]),
    'pps_gather_max_fwd': (i32, [c_f32p, c_i32p, i64, i64, i64, i32, i32, c_f32p, c_i32p, c_voidp]),
    'pps_gather_max_bwd': (i32, [c_f32p, c_i32p, i64, i32, c_f32p, c_voidp]),
    'pps_attn_pool_fwd': (i32, [c_f32p, c_f32p, i64, i32, i32, i32, c_f32p, c_f32p, c_f32p, c_voidp]),
    'pps_attn_pool_bwd': (i32, [c_f32p, c_f32p, c_f32p, c_f32p, i64, i32, i32, i32, c_f32p, c_f32p, c_voidp]),
    'pps_fka_geometry_fwd': (i32, [c_f32p, c_f32p, c_i32p, i64, i64, i64, i32, c_f32p, c_f32p, c_f32p, ctypes.c_float, i32, c_f32p, c_f32p,
                                   c_f32p, c_f32p, c_voidp, c_voidp]),
    'pps_fka_weights_bwd': (i32, [c_f32p, c_f32p, c_f32p, i64, i32, c_voidp, c_voidp]),
    'pps_fka_feat_fwd': (i32, [c_f32p, c_i32p, c_f32p, i64, i64, i64, i32, i32, c_f32p, c_voidp]),
    'pps_fka_feat_bwd': (i32, [c_f32p, c_f32p, c_i32p, c_f32p, i64, i64, i64, i32, i32, c_f32p, c_f32p, c_voidp]),
    'pps_ce_fwd': (i32, [c_f32p, c_voidp, i64, i32, c_f32p, c_voidp, c_voidp]),
    'pps_ce_bwd': (i32, [c_f32p, c_voidp, c_f32p, i64, i32, c_f32p, c_voidp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            '{} is missing. Build it with `python ppsurf_b200/build.py` (needs nvcc). '
            'ppsurf_b200 has no CPU or PyTorch fallback.'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def source_hash_of_tree() -> int:
    """digest of the CUDA sources next to this file, computed like ``ppsurf_b200/build.py`` does at build time"""
    import importlib.util
    spec = importlib.util.spec_from_file_location('ppsurf_b200_build', os.path.join(HERE, 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.source_hash()


def assert_binary_matches_sources():
    """raises when the loaded library was built from other sources than the ones in the tree (stale prebuilt .so)"""
    built, tree = int(lib.pps_source_hash()), source_hash_of_tree()
    if built != tree:
        raise PpsError('libppsurf_b200.so is stale: built from sources with digest {:#x}, the tree has {:#x}; '
                       'run `python ppsurf_b200/build.py`'.format(built, tree))


def check(status: int):
    if status != 0:
        msg = lib.pps_last_error()
        raise PpsError('ppsurf_b200 call failed ({}): {}'.format(status, msg.decode() if msg else ''))
