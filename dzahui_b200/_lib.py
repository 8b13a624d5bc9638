"""ctypes loader for libdzahui_b200.so (the C ABI of include/dzahui_b200.h).

The product has no CPU path: if the shared library is missing this raises, and every compute entry point returns
DZB_CUDA (-> CudaError) when no GPU is usable."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("DZB_LIB") or os.path.join(_HERE, "libdzahui_b200.so")  # DZB_LIB: an alternative build of the same library

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
c_size_p = C.POINTER(C.c_size_t)
FORCE_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


class EulerOp(C.Structure):
    _fields_ = [("opcode", C.c_int32), ("arg", C.c_int32), ("value", C.c_double)]


class Options(C.Structure):
    _fields_ = [("assembly_mode", C.c_int), ("solve_mode", C.c_int), ("refine_steps", C.c_int), ("stored_operators", C.c_int)]


# name -> (restype, argtypes); must list every symbol include/dzahui_b200.h declares (tests/test_abi.py checks it)
SIGNATURES = {
    "dzb_last_error": (C.c_char_p, []),
    "dzb_version": (C.c_int, []),
    "dzb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "dzb_set_device": (C.c_int, [C.c_int]),
    "dzb_set_stream": (C.c_int, [C.c_void_p]),
    "dzb_synchronize": (C.c_int, []),
    "dzb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "dzb_host_free": (C.c_int, [C.c_void_p]),
    "dzb_launch_count": (C.c_int, [C.POINTER(C.c_ulonglong), C.c_int]),
    "dzb_mesh_ingest_obj_1d": (C.c_int, [C.c_char_p, C.c_int, C.c_double, C.POINTER(C.c_void_p)]),
    "dzb_mesh_from_nodes": (C.c_int, [c_double_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "dzb_mesh_set_nodes": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_mesh_node_count": (C.c_int, [C.c_void_p, c_size_p]),
    "dzb_mesh_nodes": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_mesh_vertices": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_mesh_indices": (C.c_int, [C.c_void_p, c_u32_p, C.c_size_t]),
    "dzb_mesh_elements": (C.c_int, [C.c_void_p, c_u32_p, C.c_size_t]),
    "dzb_mesh_info": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_float_p]),
    "dzb_mesh_device_nodes": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dzb_mesh_update_gradient_1d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_float_p, C.c_size_t]),
    "dzb_mesh_destroy": (None, [C.c_void_p]),
    "dzb_mesh_ingest_obj_2d": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "dzb_mesh2d_counts": (C.c_int, [C.c_void_p, c_size_p, c_size_p, c_size_p]),
    "dzb_mesh2d_vertices": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_mesh2d_indices": (C.c_int, [C.c_void_p, c_u32_p, C.c_size_t]),
    "dzb_mesh2d_boundary_indices": (C.c_int, [C.c_void_p, c_u32_p, C.c_size_t]),
    "dzb_mesh2d_info": (C.c_int, [C.c_void_p, c_double_p, c_float_p]),
    "dzb_mesh2d_destroy": (None, [C.c_void_p]),
    "dzb_boundary_vertices_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, c_size_p]),
    "dzb_quad_pair": (C.c_int, [C.c_size_t, C.c_size_t, c_double_p, c_double_p]),
    "dzb_quad_table": (C.c_int, [C.c_size_t, c_double_p, c_double_p]),
    "dzb_stokes1d_create": (C.c_int, [C.c_void_p, C.c_double, C.c_double, FORCE_FN, C.c_void_p, C.c_size_t,
                                      C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "dzb_diffusion_ti_create": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_size_t,
                                          C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "dzb_diffusion_td_create": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, c_double_p,
                                          C.c_size_t, C.c_size_t, C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "dzb_solver_rebuild": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dzb_solver_solve": (C.c_int, [C.c_void_p, C.c_double, c_double_p, C.c_size_t]),
    "dzb_solver_solve_device": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_void_p)]),
    "dzb_solver_node_count": (C.c_int, [C.c_void_p, c_size_p]),
    "dzb_solver_path": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "dzb_solver_bands": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_size_t]),
    "dzb_solver_state": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_solver_destroy": (None, [C.c_void_p]),
    "dzb_diffusion_ti_create_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_size_t, C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "dzb_solver_block_reduce": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "dzb_solver_block_backsolve": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]),
    "dzb_solver_block_edges": (C.c_int, [C.c_void_p, c_double_p]),
    "dzb_solver_block_residual": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "dzb_solver_block_read": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_interface_solve": (C.c_int, [c_double_p, C.c_size_t, c_double_p]),
    "dzb_solver_block_reduce_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "dzb_interface_solve_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "dzb_solver_block_backsolve_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "dzb_solver_block_edges_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dzb_solver_block_residual_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dzb_solver_block_device_solution": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dzb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "dzb_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "dzb_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dzb_comm_all_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dzb_comm_destroy": (None, [C.c_void_p]),
    "dzb_solver_block_rebuild": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dzb_solver_block_solve_device": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dzb_solver_block_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dzb_format_f64": (C.c_int, [C.c_double, C.c_char_p, C.c_size_t]),
    "dzb_csv_prepare_dir": (C.c_int, [C.c_char_p, C.c_int]),
    "dzb_csv_write": (C.c_int, [C.c_char_p, C.c_char_p, C.c_double, C.POINTER(C.c_char_p), C.c_size_t, c_double_p, C.c_size_t]),
    "dzb_ensemble_td_create": (C.c_int, [c_double_p, C.c_size_t, C.c_size_t, C.c_double, C.c_double, C.c_double,
                                         C.c_double, c_double_p, C.c_size_t, C.c_size_t, C.POINTER(Options),
                                         C.POINTER(C.c_void_p)]),
    "dzb_ensemble_step": (C.c_int, [C.c_void_p, C.c_double, C.c_size_t]),
    "dzb_ensemble_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]),
    "dzb_ensemble_solve_f32": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]),
    "dzb_ensemble_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dzb_ensemble_read_bars": (C.c_int, [C.c_void_p, c_size_p, C.c_size_t, c_double_p]),
    "dzb_ensemble_state_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dzb_ensemble_bands": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, c_double_p, c_double_p, c_double_p, C.c_size_t]),
    "dzb_ensemble_dims": (C.c_int, [C.c_void_p, c_size_p, c_size_p]),
    "dzb_ensemble_step_kernel": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "dzb_ensemble_destroy": (None, [C.c_void_p]),
    "dzb_euler_create": (C.c_int, [C.c_size_t, C.c_size_t, c_double_p, c_double_p, C.POINTER(C.c_void_p)]),
    "dzb_euler_create_program": (C.c_int, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, c_double_p, C.c_size_t, c_double_p,
                                           C.POINTER(C.c_void_p)]),
    "dzb_euler_set_values": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_euler_step": (C.c_int, [C.c_void_p, C.c_double, C.c_size_t]),
    "dzb_euler_values": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t]),
    "dzb_euler_values_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dzb_euler_destroy": (None, [C.c_void_p]),
}

_lib = None


def load():
    """Load the C-ABI library (built in-tree by dzahui_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m dzahui_b200.build` (nvcc, sm_100a). "
            "dzahui_b200 has no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
