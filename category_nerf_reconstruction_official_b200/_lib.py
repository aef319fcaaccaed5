"""ctypes binding of the C-ABI library ``csrc/libcnf_sm100.so`` (declared in ``include/cnf.h``).

The library is built in-tree with nvcc for sm_100a (``build()``), loaded with ctypes and called with raw
device pointers + the current CUDA stream.  There is NO CPU fallback: every op raises if the library is
missing or the device is not an sm_100 part.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libcnf_sm100.so")
# development build: the same sources + cnf_selftest.cu with -DCNF_DEV (include/cnf_dev.h: descriptor self-tests, probes, clock stamps)
DEV_LIB_PATH = os.path.join(CSRC, "libcnf_sm100_dev.so")
SOURCES = ["cnf_api.cu", "cnf_sampler.cu", "cnf_embed.cu", "cnf_composite.cu", "cnf_mlp_simt.cu", "cnf_fused.cu", "cnf_occmap_tc.cu"]
DEV_SOURCES = SOURCES + ["cnf_selftest.cu"]
HEADERS = ["cnf_common.cuh", "cnf_umma.cuh", "cnf_sample_math.cuh", "cnf_fused_a.inc", "cnf_fused_b.inc", "cnf_fused_c.inc", os.path.join("..", "..", "include", "cnf.h"),
           os.path.join("..", "..", "include", "cnf_dev.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]

_lock = threading.Lock()
_lib = None
_dev_lib = None


class CnfError(RuntimeError):
    pass


def _stale(lib_path=None, sources=None) -> bool:
    lib_path = lib_path or LIB_PATH
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    for f in (sources or SOURCES) + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def _compile(lib_path, sources, extra, tag, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    for src in sources:
        obj = os.path.join(CSRC, "build", src.replace(".cu", tag + ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise CnfError("nvcc failed for %s:\n%s" % (src, out))
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", lib_path] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise CnfError("link failed:\n" + r.stdout)
    return lib_path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libcnf_sm100.so in-tree (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    return _compile(LIB_PATH, SOURCES, [], "", verbose)


def build_dev(force: bool = False, verbose: bool = False) -> str:
    """The development build (-DCNF_DEV): product ABI + include/cnf_dev.h.  Used by the hardware self-tests and tuning scripts."""
    if not force and not _stale(DEV_LIB_PATH, DEV_SOURCES):
        return DEV_LIB_PATH
    return _compile(DEV_LIB_PATH, DEV_SOURCES, ["-DCNF_DEV"], ".dev", verbose)


# name -> (restype, argtypes); must list every symbol declared in include/cnf.h
P = c_void_p
SIGNATURES = {
    "cnf_version": (c_int, []),
    "cnf_last_error_string": (c_char_p, []),
    "cnf_device_check": (c_int, [c_int]),
    "cnf_sample_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cnf_sample_fwd": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P,
                               c_float, c_float, c_float, c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "cnf_sample_stats": (c_int, [c_int, c_int, P, P, c_float, c_int, P, P, c_size_t, P]),
    "cnf_reduce_stats": (c_int, [c_int, c_int, P, P, P, P]),
    "cnf_stratified_bins": (c_int, [c_int, c_int, P, c_int, P, c_int, P, P, P, P]),
    "cnf_pool_gather": (c_int, [c_int, c_int, c_int] + [P] * 15),
    "cnf_pool_compose": (c_int, [ctypes.c_int64, P, P, P, P]),
    "cnf_embed_fwd": (c_int, [c_int, c_int, P, P, P, P, P]),
    "cnf_embed_bwd": (c_int, [c_int, c_int, P, P, P, P, P, P, P]),
    "cnf_codenerf_param_count": (c_size_t, [c_int, c_int]),
    "cnf_codenerf_workspace_floats": (c_size_t, [c_int, c_int, c_int]),
    "cnf_codenerf_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "cnf_codenerf_bwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_size_t,
                                 P, P, P, P, P]),
    "cnf_occmap_param_count": (c_size_t, [c_int]),
    "cnf_occmap_workspace_floats": (c_size_t, [c_int, c_int]),
    "cnf_occmap_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, c_size_t, P]),
    "cnf_occmap_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, c_size_t, P, P, P]),
    "cnf_termination_fwd": (c_int, [c_int, c_int, P, P, P]),
    "cnf_termination_bwd": (c_int, [c_int, c_int, P, P, P, P]),
    "cnf_ray_uncertainty": (c_int, [c_int, c_int, P, P, P]),
    "cnf_mask_counts": (c_int, [c_int, c_int, P, P, P, P]),
    "cnf_composite_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cnf_composite_loss_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_float, c_float, P, P, P,
                                       c_size_t, P]),
    "cnf_composite_loss_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_float, c_float, P, P, P, P]),
    "cnf_code_reg": (c_int, [c_int, c_int, P, c_float, P, P, P]),
    "cnf_train_step_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "cnf_train_step": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P, P, P,
                               c_float, c_float, c_float, c_int, P, P, P, P, P, P, c_size_t, P]),
    "cnf_train_step_sampled_max_rays": (c_int, []),
    "cnf_train_step_sampled": (c_int, [c_int] * 6 + [P] * 18 + [c_float, c_float, c_float, c_int, P, c_float, c_float, c_float, c_int]
                               + [P] * 8 + [c_size_t, P]),
    "cnf_train_step_prepare": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, P, c_size_t, P]),
    "cnf_eval_points_workspace_bytes": (c_size_t, []),
    "cnf_eval_points": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "cnf_render_rays_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cnf_render_rays": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "cnf_composite_loss_fwd_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_float, c_float, P, P, P, P, P, c_size_t, P]),
    "cnf_origin_dirs": (c_int, [c_int, c_int, P, P, P, P, P]),
    "cnf_occmap_tc_workspace_bytes": (c_size_t, [c_int]),
    "cnf_occmap_tc_fwd": (c_int, [c_int, P, P, P, P, P, c_size_t, P]),
    "cnf_occmap_tc_fwd_points": (c_int, [c_int, P, P, P, P, P, P, P, c_size_t, c_int, P]),
    "cnf_occmap_tc_bwd_points": (c_int, [c_int, P, P, P, P, P, P, c_size_t, P, P, c_int, P]),
    "cnf_occmap_tc_prepare": (c_int, [c_int, P, P, P, P, c_size_t, P]),
    "cnf_occmap_tc_eval_workspace_bytes": (c_size_t, []),
    "cnf_occmap_tc_eval_points": (c_int, [c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "cnf_occmap_tc_render_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cnf_occmap_tc_render_rays": (c_int, [c_int, c_int, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "cnf_occmap_tc_bwd": (c_int, [c_int, P, P, P, P, P, c_size_t, P, P, P]),
    "cnf_set_kernel_events": (c_int, [P, P]),
    "cnf_adamw_step": (c_int, [c_size_t, P, P, P, P, c_float, c_float, c_float, c_float, c_float, c_int, P]),
    "cnf_adamw_groups": (c_int, [c_int, P, c_float, c_float, c_float, P, P]),
}


# include/cnf_dev.h (development build only)
DEV_SIGNATURES = {
    "cnf_occmap_tc_set_debug": (c_int, [P]),
    "cnf_set_debug_buffer": (c_int, [P]),
    "cnf_umma_m64_probe": (c_int, [c_int, c_int, c_int, P, P, P, P]),
    "cnf_umma_rate_probe": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "cnf_umma_selftest": (c_int, [c_int, c_int, c_int, P, P, P, P]),
}


class AdamwGroup(ctypes.Structure):
    """include/cnf.h: cnf_adamw_group"""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("n", c_size_t),
                ("lr", c_float), ("weight_decay", c_float)]


def load(auto_build: bool = True):
    """dlopen the library (building it first if sources are newer). Raises CnfError if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("CNF_LIB_PATH", LIB_PATH)      # development aid: load a variant build of the same library
        if path == LIB_PATH and auto_build and _stale():
            build()
        if not os.path.exists(path):
            raise CnfError("libcnf_sm100.so is missing (%s): run __graft_entry__.build(); there is no fallback path"
                           % path)
        _lib = _bind(ctypes.CDLL(path))
        return _lib


def _bind(lib, dev: bool = False):
    sigs = dict(SIGNATURES)
    if dev or hasattr(lib, "cnf_set_debug_buffer"):       # a development build loaded through CNF_LIB_PATH
        sigs.update(DEV_SIGNATURES)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)       # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


def load_dev(auto_build: bool = True):
    """dlopen the development build (product ABI + include/cnf_dev.h); built on demand."""
    global _dev_lib
    with _lock:
        if _dev_lib is None:
            if auto_build and _stale(DEV_LIB_PATH, DEV_SOURCES):
                build_dev()
            if not os.path.exists(DEV_LIB_PATH):
                raise CnfError("libcnf_sm100_dev.so is missing: run category_nerf_reconstruction_official_b200.build_dev()")
            _dev_lib = _bind(ctypes.CDLL(DEV_LIB_PATH), dev=True)
        return _dev_lib


_checked_devices = set()


def lib_for(t: torch.Tensor):
    """Library handle for an op on tensor ``t``; enforces CUDA + sm_100."""
    if not t.is_cuda:
        raise CnfError("cnf ops run only on CUDA (sm_100a) tensors; got device %s. There is no CPU fallback." % t.device)
    lib = load()
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _checked_devices:
        check(lib.cnf_device_check(dev))
        _checked_devices.add(dev)
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().cnf_last_error_string()
        raise CnfError("libcnf error %d: %s" % (rc, msg.decode() if msg else "?"))


def ptr(t):
    """data pointer of a contiguous tensor (or NULL for None)."""
    if t is None:
        return None
    assert t.is_contiguous(), "cnf: tensor must be contiguous"
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
