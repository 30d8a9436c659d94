"""ctypes binding of libkmos_b200.so (include/kmos_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no GPU is visible, stepping calls
raise.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkmos_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

OK = 0
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_SMEM, KERNEL_WARP_HBM, KERNEL_GENERATED, KERNEL_OTF_FAST = 0, 1, 2, 3, 4, 5
BACKEND_LOCAL_SMART, BACKEND_LAT_INT, BACKEND_OTF = 0, 1, 2
REPLICA_OK, REPLICA_DEADLOCK, REPLICA_SPECIES_MISMATCH, REPLICA_CAPACITY, REPLICA_BAD_MODEL = range(5)


class KmosB200Error(RuntimeError):
    pass


def sources():
    return [os.path.join(CSRC, f) for f in ("kmos_b200.cu", "kb_smem.cuh", "kb_latint.cuh", "kb_otf.cuh", "kb_otf_event.cuh", "kb_otf_fast.cuh", "kb_gen.cuh", "kb_fleet.h",
                                            "kb_interp.h", "kb_common.h")] + \
        [os.path.join(os.path.dirname(HERE), "include", "kmos_b200.h")]


STAMP = LIB + ".sha256"


def source_digest():
    """Content hash of everything the library is built from (sources + flags): the staleness key."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for s in sources():
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_current():
    """True if the built library matches the sources byte for byte (a checkout that resets mtimes cannot fool it)."""
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_digest()


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> kmos_b200/libkmos_b200.so (in-tree)."""
    srcs = sources()
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, srcs[0], "-ldl"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(source_digest() + "\n")
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise KmosB200Error("%s not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB)
    if not is_current():
        # a library older than its sources is never used silently: rebuild it, or say so
        import shutil
        if shutil.which(os.environ.get("NVCC", "nvcc")):
            build()
        else:
            raise KmosB200Error("%s does not match its sources and there is no nvcc to rebuild it" % LIB)
    L = C.CDLL(LIB)
    vp, i32, i64, u32, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double

    def arr(dt):
        return np.ctypeslib.ndpointer(dtype=dt, flags="C_CONTIGUOUS")

    sig = {
        "kmos_b200_last_error": (C.c_char_p, []),
        "kmos_b200_device_count": (C.c_int, []),
        "kmos_b200_model_create": (C.c_int, [arr(np.int32), i64, C.POINTER(vp)]),
        "kmos_b200_model_destroy": (None, [vp]),
        "kmos_b200_model_nproc": (C.c_int, [vp]),
        "kmos_b200_model_nspecies": (C.c_int, [vp]),
        "kmos_b200_model_spuck": (C.c_int, [vp]),
        "kmos_b200_model_lut_size": (C.c_int, [vp]),
        "kmos_b200_batch_create": (C.c_int, [vp, i32, arr(np.int32), i32, C.POINTER(vp)]),
        "kmos_b200_batch_destroy": (None, [vp]),
        "kmos_b200_batch_volume": (C.c_int, [vp]),
        "kmos_b200_select_kernel": (C.c_int, [vp, i32]),
        "kmos_b200_batch_attach_proclist": (C.c_int, [vp, C.c_char_p]),
        "kmos_b200_batch_detach_proclist": (C.c_int, [vp]),
        "kmos_b200_kernel_info": (C.c_int, [vp, arr(np.int64)]),
        "kmos_b200_set_seeds": (C.c_int, [vp, arr(np.uint64), vp]),
        "kmos_b200_set_rates": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_set_rate_const": (C.c_int, [vp, i32, i32, f64]),
        "kmos_b200_get_rates": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_set_otf_lut": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_init_state": (C.c_int, [vp, i32]),
        "kmos_b200_set_configuration": (C.c_int, [vp, i32, arr(np.int32), i32]),
        "kmos_b200_do_kmc_steps": (C.c_int, [vp, i64]),
        "kmos_b200_synchronize": (C.c_int, [vp]),
        "kmos_b200_reload_replica": (C.c_int, [vp, i32, arr(np.int32), arr(np.int32), arr(np.int32), arr(np.int64),
                                               arr(np.float64), f64, i64]),
        "kmos_b200_get_next_kmc_step": (C.c_int, [vp, arr(np.int32), arr(np.int32)]),
        "kmos_b200_run_proc_nr": (C.c_int, [vp, arr(np.int32), arr(np.int32)]),
        "kmos_b200_timer_start": (C.c_int, [vp]),
        "kmos_b200_timer_stop": (C.c_int, [vp, C.POINTER(f64)]),
        "kmos_b200_get_kmc_time": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_get_kmc_time_step": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_get_kmc_step": (C.c_int, [vp, arr(np.int64)]),
        "kmos_b200_set_kmc_time": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_get_procstat": (C.c_int, [vp, arr(np.int64)]),
        "kmos_b200_get_integ_rates": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_get_nr_of_sites": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_get_accum_rates": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_get_lattice": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_get_occupation": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_get_avail_sites": (C.c_int, [vp, i32, arr(np.int32)]),
        "kmos_b200_get_status": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_get_error_info": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_tally_words": (C.c_int, [vp]),
        "kmos_b200_reduce_tallies": (C.c_int, [vp, vp, i32, vp, vp]),
        "kmos_b200_philox_next": (f64, [u64, u32, u64, i32]),
        "kmos_b200_batch_set_stream": (C.c_int, [vp, vp]),
        "kmos_b200_measure_smem_bandwidth": (C.c_int, [i32, C.POINTER(f64), C.POINTER(f64)]),
        # fleet: the same replicas on several GPUs of this process (csrc/kb_fleet.h)
        "kmos_b200_fleet_create": (C.c_int, [vp, i32, arr(np.int32), vp, arr(np.int32), i32, C.POINTER(vp)]),
        "kmos_b200_fleet_destroy": (None, [vp]),
        "kmos_b200_fleet_n_shards": (C.c_int, [vp]),
        "kmos_b200_fleet_shard": (vp, [vp, i32, C.POINTER(i32), C.POINTER(i32)]),
        "kmos_b200_fleet_attach_proclist": (C.c_int, [vp, C.c_char_p]),
        "kmos_b200_fleet_select_kernel": (C.c_int, [vp, i32]),
        "kmos_b200_fleet_set_rates": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_fleet_set_otf_lut": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_fleet_init_state": (C.c_int, [vp, i32]),
        "kmos_b200_fleet_do_kmc_steps": (C.c_int, [vp, i64]),
        "kmos_b200_fleet_synchronize": (C.c_int, [vp]),
        "kmos_b200_fleet_get_kmc_time": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_fleet_get_kmc_step": (C.c_int, [vp, arr(np.int64)]),
        "kmos_b200_fleet_get_status": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_fleet_get_procstat": (C.c_int, [vp, arr(np.int64)]),
        "kmos_b200_fleet_get_integ_rates": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_fleet_get_nr_of_sites": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_fleet_get_lattice": (C.c_int, [vp, arr(np.int32)]),
        "kmos_b200_fleet_get_occupation": (C.c_int, [vp, arr(np.float64)]),
        "kmos_b200_fleet_reduce_tallies": (C.c_int, [vp, vp, i32, arr(np.float64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


EXPORTED = [
    "kmos_b200_last_error", "kmos_b200_device_count", "kmos_b200_model_create", "kmos_b200_model_destroy",
    "kmos_b200_model_nproc", "kmos_b200_model_nspecies", "kmos_b200_model_spuck", "kmos_b200_model_lut_size",
    "kmos_b200_batch_create", "kmos_b200_batch_destroy", "kmos_b200_batch_volume", "kmos_b200_select_kernel",
    "kmos_b200_kernel_info", "kmos_b200_set_seeds", "kmos_b200_set_rates", "kmos_b200_set_rate_const",
    "kmos_b200_get_rates", "kmos_b200_set_otf_lut", "kmos_b200_init_state", "kmos_b200_set_configuration",
    "kmos_b200_do_kmc_steps", "kmos_b200_synchronize", "kmos_b200_timer_start", "kmos_b200_timer_stop",
    "kmos_b200_get_kmc_time", "kmos_b200_get_kmc_time_step", "kmos_b200_get_kmc_step", "kmos_b200_set_kmc_time",
    "kmos_b200_get_procstat", "kmos_b200_get_integ_rates", "kmos_b200_get_nr_of_sites",
    "kmos_b200_get_accum_rates", "kmos_b200_get_lattice", "kmos_b200_get_occupation",
    "kmos_b200_get_avail_sites", "kmos_b200_get_status", "kmos_b200_get_error_info", "kmos_b200_tally_words",
    "kmos_b200_reduce_tallies", "kmos_b200_philox_next", "kmos_b200_batch_set_stream",
    "kmos_b200_measure_smem_bandwidth", "kmos_b200_get_next_kmc_step", "kmos_b200_run_proc_nr",
    "kmos_b200_reload_replica", "kmos_b200_batch_attach_proclist", "kmos_b200_batch_detach_proclist",
    "kmos_b200_fleet_create", "kmos_b200_fleet_destroy", "kmos_b200_fleet_n_shards", "kmos_b200_fleet_shard",
    "kmos_b200_fleet_attach_proclist", "kmos_b200_fleet_select_kernel", "kmos_b200_fleet_set_rates",
    "kmos_b200_fleet_set_otf_lut", "kmos_b200_fleet_init_state", "kmos_b200_fleet_do_kmc_steps",
    "kmos_b200_fleet_synchronize", "kmos_b200_fleet_get_kmc_time", "kmos_b200_fleet_get_kmc_step",
    "kmos_b200_fleet_get_status", "kmos_b200_fleet_get_procstat", "kmos_b200_fleet_get_integ_rates",
    "kmos_b200_fleet_get_nr_of_sites", "kmos_b200_fleet_get_lattice", "kmos_b200_fleet_get_occupation",
    "kmos_b200_fleet_reduce_tallies",
]


def check(rc):
    if rc != OK:
        raise KmosB200Error("kmos_b200 error %d: %s" % (rc, lib().kmos_b200_last_error().decode()))
