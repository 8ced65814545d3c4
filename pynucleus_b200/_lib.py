"""ctypes binding of libpnb200.so (C ABI: include/pnb200.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``pynucleus_b200.build.build()``.  There is no CPU fallback: if the shared
library is missing, or no CUDA device is visible, every compute call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpnb200.so')

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class pnb_mesh_t(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int32), ('num_vertices', ctypes.c_int32), ('num_cells', ctypes.c_int32),
                ('vertices', ctypes.c_void_p), ('cells', ctypes.c_void_p), ('vol', ctypes.c_void_p),
                ('h', ctypes.c_void_p), ('diam', ctypes.c_double), ('num_bfacets', ctypes.c_int32),
                ('bfacets', ctypes.c_void_p), ('num_blocks', ctypes.c_int32), ('block_cell_ptr', ctypes.c_void_p),
                ('block_dof_ptr', ctypes.c_void_p), ('block_facet_ptr', ctypes.c_void_p)]


class pnb_dofmap_t(ctypes.Structure):
    _fields_ = [('dofs_per_element', ctypes.c_int32), ('num_dofs', ctypes.c_int32), ('dofs', ctypes.c_void_p)]


class pnb_kernel_t(ctypes.Structure):
    _fields_ = [('kernel_type', ctypes.c_int32), ('dim', ctypes.c_int32), ('s', ctypes.c_double),
                ('scaling', ctypes.c_double), ('bscaling', ctypes.c_double), ('singularity', ctypes.c_double),
                ('bsingularity', ctypes.c_double), ('horizon2', ctypes.c_double),
                ('target_order', ctypes.c_double), ('btarget_order', ctypes.c_double), ('order_num_dofs', ctypes.c_int32),
                ('cell_labels', ctypes.c_void_p), ('bfacet_labels', ctypes.c_void_p), ('active_class', ctypes.c_int32),
                ('pair_class', ctypes.c_uint8*16), ('bpair_class', ctypes.c_uint8*16), ('pair_orientation', ctypes.c_int32), ('pair_filter', ctypes.c_int32)]


class pnb_rule_t(ctypes.Structure):
    _fields_ = [('n', ctypes.c_int32), ('rows', ctypes.c_int32), ('bary', ctypes.c_void_p), ('w', ctypes.c_void_p)]


class pnb_rules_t(ctypes.Structure):
    _fields_ = [('identical', pnb_rule_t), ('edge', pnb_rule_t), ('vertex', pnb_rule_t),
                ('bedge', pnb_rule_t), ('bvertex', pnb_rule_t), ('max_order', ctypes.c_int32),
                ('cell', ctypes.POINTER(pnb_rule_t)), ('facet', ctypes.POINTER(pnb_rule_t))]


class pnb_varorder_t(ctypes.Structure):
    _fields_ = [('fun', ctypes.c_int32), ('sl', ctypes.c_double), ('sr', ctypes.c_double), ('r', ctypes.c_double),
                ('slope', ctypes.c_double), ('interface', ctypes.c_double), ('num_values', ctypes.c_int32),
                ('values', ctypes.c_void_p), ('cell_value', ctypes.c_void_p), ('bfacet_value', ctypes.c_void_p),
                ('identical', ctypes.POINTER(pnb_rule_t)), ('edge', ctypes.POINTER(pnb_rule_t)),
                ('vertex', ctypes.POINTER(pnb_rule_t)), ('bedge', ctypes.POINTER(pnb_rule_t)),
                ('bvertex', ctypes.POINTER(pnb_rule_t)), ('vertex_values', ctypes.c_void_p)]


class pnb_h2_desc_t(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int32), ('num_dofs', ctypes.c_int32), ('num_nodes', ctypes.c_int32),
                ('coef_ptr', ctypes.c_void_p), ('parent', ctypes.c_void_p), ('level', ctypes.c_void_p),
                ('num_leaves', ctypes.c_int32), ('leaf_node', ctypes.c_void_p), ('leaf_dof_ptr', ctypes.c_void_p),
                ('leaf_dofs', ctypes.c_void_p), ('leaf_values', ctypes.c_void_p), ('leaf_cell_ptr', ctypes.c_void_p),
                ('leaf_cells', ctypes.c_void_p), ('leaf_cell_pos', ctypes.c_void_p), ('leaf_boxes', ctypes.c_void_p),
                ('leaf_orders', ctypes.c_void_p), ('num_vertices', ctypes.c_int32), ('num_cells', ctypes.c_int32),
                ('vertices', ctypes.c_void_p), ('cells', ctypes.c_void_p), ('vol', ctypes.c_void_p),
                ('max_m', ctypes.c_int32), ('rule_n', ctypes.c_void_p), ('rule_bary_ptr', ctypes.c_void_p),
                ('rule_w_ptr', ctypes.c_void_p), ('rule_bary', ctypes.c_void_p), ('rule_w', ctypes.c_void_p),
                ('rule_bary_size', ctypes.c_int64), ('rule_w_size', ctypes.c_int64),
                ('eta', ctypes.c_void_p), ('eta_ptr', ctypes.c_void_p),
                ('transfer_ptr', ctypes.c_void_p), ('transfer', ctypes.c_void_p), ('transfer_size', ctypes.c_int64),
                ('num_far', ctypes.c_int32), ('far_n1', ctypes.c_void_p), ('far_n2', ctypes.c_void_p),
                ('far_ptr', ctypes.c_void_p), ('far_blocks', ctypes.c_void_p), ('far_size', ctypes.c_int64),
                ('near_indptr', ctypes.c_void_p), ('near_indices', ctypes.c_void_p), ('near_data', ctypes.c_void_p)]


# every symbol include/pnb200.h declares
EXPORTS = ['pnb_last_error', 'pnb_version', 'pnb_device_count', 'pnb_problem_create', 'pnb_problem_set_rules',
           'pnb_problem_destroy', 'pnb_max_order', 'pnb_classify_pairs', 'pnb_panel_histogram',
           'pnb_local_matrices', 'pnb_far_max_order', 'pnb_dense_assemble', 'pnb_dense_stats',
           'pnb_dense_timings', 'pnb_dense_matvec', 'pnb_fp64_peak', 'pnb_row_granularity', 'pnb_dense_rows_begin',
           'pnb_dense_cell_blocks', 'pnb_dense_cell_blocks_copy', 'pnb_dense_rows_end', 'pnb_farfield_blocks',
           'pnb_release_cached_memory', 'pnb_dense_kernel_timings', 'pnb_dist_plan', 'pnb_dist_rows', 'pnb_dist_eval',
           'pnb_dist_status', 'pnb_dist_apply', 'pnb_device_alloc', 'pnb_device_free', 'pnb_ipc_export', 'pnb_ipc_import',
           'pnb_ipc_close',
           'pnb_boundary_cell_blocks', 'pnb_mesh_edge_lengths', 'pnb_problem_set_path', 'pnb_sparsity_mask', 'pnb_block_alignment',
           'pnb_h2_create', 'pnb_h2_leaf_values', 'pnb_h2_matvec', 'pnb_h2_destroy', 'pnb_dense_assemble_element', 'pnb_dense_assemble_element_tempered', 'pnb_dense_assemble_element_smooth', 'pnb_dense_assemble_varorder', 'pnb_problem_set_row_part', 'pnb_element_rows', 'pnb_element_rows_host',
           'pnb_krylov_workspace_doubles', 'pnb_krylov_dot', 'pnb_krylov_cg_update', 'pnb_krylov_cg_direction']

_LIB = None


class PNBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libpnb200 error {}: {}'.format(code, msg))
        self.code = code


def lib():
    """Load libpnb200.so; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('{} not found: build it with `python -c "import __graft_entry__ as g; g.build()"`. '
                               'There is no CPU fallback.'.format(LIB_PATH))
        L = ctypes.CDLL(LIB_PATH)
        L.pnb_last_error.restype = ctypes.c_char_p
        L.pnb_problem_create.argtypes = [ctypes.POINTER(pnb_mesh_t), ctypes.POINTER(pnb_dofmap_t),
                                         ctypes.POINTER(pnb_kernel_t), ctypes.POINTER(pnb_rules_t),
                                         ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.pnb_problem_set_rules.argtypes = [ctypes.c_void_p, ctypes.POINTER(pnb_rules_t)]
        L.pnb_problem_destroy.argtypes = [ctypes.c_void_p]
        L.pnb_problem_set_path.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.pnb_sparsity_mask.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.pnb_problem_destroy.restype = None
        L.pnb_max_order.argtypes = [ctypes.c_void_p, ctypes.c_int, c_int32_p]
        L.pnb_classify_pairs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_panel_histogram.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_local_matrices.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_dense_assemble.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.pnb_dense_rows_begin.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_void_p, ctypes.c_int64]
        L.pnb_dist_plan.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_int32_p, c_int64_p]
        L.pnb_dist_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_dist_eval.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.pnb_dist_status.argtypes = [ctypes.c_void_p, c_int32_p]
        L.pnb_dist_apply.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        L.pnb_device_alloc.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]
        L.pnb_device_free.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.pnb_ipc_export.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_ipc_import.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        L.pnb_ipc_close.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.pnb_dense_rows_end.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]
        L.pnb_dense_cell_blocks.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), c_int64_p]
        L.pnb_dense_cell_blocks_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.pnb_boundary_cell_blocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_farfield_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]
        L.pnb_dense_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_dense_timings.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_dense_kernel_timings.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_dense_matvec.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_fp64_peak.argtypes = [ctypes.c_int, c_double_p]
        L.pnb_mesh_edge_lengths.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            c_double_p, c_double_p]
        L.pnb_h2_create.argtypes = [ctypes.c_int, ctypes.POINTER(pnb_h2_desc_t), ctypes.POINTER(ctypes.c_void_p)]
        L.pnb_h2_leaf_values.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_h2_matvec.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.pnb_h2_destroy.argtypes = [ctypes.c_void_p]
        L.pnb_dense_assemble_element.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.pnb_dense_assemble_element_tempered.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.pnb_dense_assemble_element_smooth.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                        ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.pnb_problem_set_row_part.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
        L.pnb_element_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p, c_int32_p]
        L.pnb_element_rows_host.argtypes = [ctypes.c_int32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_void_p, c_int32_p]
        L.pnb_dense_assemble_varorder.argtypes = [ctypes.c_void_p, ctypes.POINTER(pnb_varorder_t), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.pnb_krylov_dot.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p]
        L.pnb_krylov_cg_update.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.pnb_krylov_cg_direction.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise PNBError(rc, lib().pnb_last_error().decode())


def as_rule(bary, w, keep):
    bary = np.ascontiguousarray(bary, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    keep += [bary, w]
    return pnb_rule_t(bary.shape[1], bary.shape[0], bary.ctypes.data, w.ctypes.data)
