"""ctypes binding of the C-ABI library declared in include/gg_raster.h.

The product path has no CPU fallback: if `csrc/libgg_raster.so` is missing and cannot be built,
or a call fails, a RuntimeError is raised."""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
import tempfile
import threading
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("GG_RASTER_LIB") or os.path.join(CSRC, "libgg_raster.so")   # override: dev experiments
SOURCES = ["project.cu", "binning.cu", "blend_fwd.cu", "blend_fwd2.cu", "blend_bwd.cu", "blend_bwd2.cu", "preprocess_bwd.cu", "mesh_binding.cu", "photometric.cu", "visibility.cu", "allreduce.cu", "c_api.cu"]
HEADERS = ["common.cuh", "mesh_binding_math.h", os.path.join("..", "..", "include", "gg_raster.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None


class GGView(C.Structure):
    _fields_ = [("num_gaussians", C.c_int32), ("sh_coeffs", C.c_int32), ("sh_degree", C.c_int32),
                ("image_width", C.c_int32), ("image_height", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("prefiltered", C.c_int32), ("debug", C.c_int32)]


class GGInputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("means3D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3D_precomp",
                 "bg", "viewmatrix", "projmatrix", "campos")]


STAMP_PATH = os.path.join(CSRC, ".build_stamp")
LOCK_PATH = os.path.join(CSRC, ".build.lock")

# project.cu is compiled with -fmad=false: without FMA contraction its IEEE +,-,*,/,sqrt arithmetic is
# bit-identical to the CPU oracle's (gcc -ffp-contract=off), so every DISCRETE decision taken from the
# projected geometry (cull, radius, tile rectangle, depth order) agrees with the oracle by construction.
PER_FILE_FLAGS = {"project.cu": ["-fmad=false"], "visibility.cu": ["-fmad=false"]}   # ray casts: bit-identical to oracle/raycast_oracle.c


def _source_digest() -> str:
    """sha256 over the sources, headers and flags: staleness is decided by CONTENT (mtimes do not survive a copy to
    another box, and a stale library must never be loaded silently)."""
    h = hashlib.sha256()
    h.update(repr((NVCC_FLAGS, sorted(PER_FILE_FLAGS.items()))).encode())
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        h.update(f.encode())
        if os.path.exists(p):
            with open(p, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def _needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or os.environ.get("GG_RASTER_LIB"):
        return not os.path.exists(LIB_PATH)
    try:
        with open(STAMP_PATH) as fh:
            return fh.read().strip() != _source_digest()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libgg_raster.so (in-tree, so it travels to the GPU box).
    Safe under torchrun: an inter-process file lock serialises builders, objects go to a private temp directory and the
    finished library is moved into place atomically -- a rank can never dlopen a half-written file."""
    if not force and not _needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("gaussian-garments_b200: nvcc not found and libgg_raster.so is missing/stale")
    import fcntl
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _needs_build():          # another process built it while we waited
                return LIB_PATH
            digest = _source_digest()
            tmp = tempfile.mkdtemp(prefix=".build_", dir=CSRC)
            try:
                common = [f for f in NVCC_FLAGS if f != "-shared"]
                procs, objs, log = [], [], ""
                for src in SOURCES:
                    obj = os.path.join(tmp, src.replace(".cu", ".o"))
                    objs.append(obj)
                    cmd = [nvcc] + common + PER_FILE_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
                    procs.append((src, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
                for src, pr in procs:
                    out, _ = pr.communicate()
                    log += out
                    if pr.returncode != 0:
                        raise RuntimeError(f"nvcc failed on {src}:\n{out}")
                tmp_lib = os.path.join(tmp, "libgg_raster.so")
                res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp_lib] + objs,
                                     cwd=CSRC, capture_output=True, text=True)
                if res.returncode != 0:
                    raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
                os.replace(tmp_lib, LIB_PATH)             # atomic on the same filesystem
                with open(STAMP_PATH + ".tmp", "w") as fh:
                    fh.write(digest)
                os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
                if verbose:
                    print(log)
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def load():
    """Load the library (building it first when sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _needs_build():
            try:
                build()
            except Exception as e:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "gaussian-garments_b200: CUDA extension libgg_raster.so is missing and could not be "
                        f"built ({e}); there is no CPU fallback") from e
                # a library exists but does not match the sources and cannot be rebuilt: never load it silently
                if os.environ.get("GG_ALLOW_STALE_LIB") != "1":
                    raise RuntimeError(
                        "gaussian-garments_b200: libgg_raster.so does not match csrc/ (content hash) and rebuilding "
                        f"failed ({e}); set GG_ALLOW_STALE_LIB=1 to load it anyway") from e
                warnings.warn(f"gaussian-garments_b200: loading a STALE libgg_raster.so (rebuild failed: {e})")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
        P = C.POINTER
        lib.gg_abi_version.restype = C.c_int
        lib.gg_version.restype = C.c_char_p
        lib.gg_last_error.restype = C.c_char_p
        lib.gg_launch_count.restype = C.c_int64
        lib.gg_launch_count.argtypes = [i32]
        lib.gg_forward_workspace_bytes.argtypes = [P(GGView), P(sz), P(sz), P(sz)]
        lib.gg_instance_workspace_bytes.argtypes = [i64, P(sz), P(sz)]
        lib.gg_backward_workspace_bytes.argtypes = [P(GGView), P(sz)]
        lib.gg_forward_project.argtypes = [P(GGView), P(GGInputs), vp, vp, vp, vp, i32, vp]
        lib.gg_forward_color.argtypes = [P(GGView), P(GGInputs), vp, vp, i32, vp]
        lib.gg_forward_render.argtypes = [P(GGView), P(GGInputs), vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, i32, vp]
        lib.gg_forward_render_late_color.argtypes = [P(GGView), P(GGInputs), vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        lib.gg_gate_signal.argtypes = [vp, i32, vp]
        lib.gg_forward_overflow_check.argtypes = [P(GGView), vp, i64, vp, i32, vp]
        lib.gg_backward.argtypes = [P(GGView), P(GGInputs), vp, vp, i64, vp, vp, vp] + [vp] * 11 + [i32, vp]
        lib.gg_mark_visible.argtypes = [C.c_int32, vp, vp, vp, vp, i32, vp]
        lib.gg_debug_read_geom.argtypes = [P(GGView), vp, vp, vp, vp, vp, vp, i32, vp]
        lib.gg_debug_lazy_phase_counters.argtypes = [vp]
        lib.gg_debug_read_binning.argtypes = [P(GGView), vp, vp, i64, vp, vp, i32, vp]
        lib.gg_mesh_bind_workspace_bytes.argtypes = [C.c_int32, P(sz)]
        lib.gg_mesh_bind_forward.argtypes = [C.c_int32] * 3 + [vp] * 10 + [i32, vp]
        lib.gg_mesh_bind_backward.argtypes = [C.c_int32] * 3 + [vp] * 15 + [i32, vp]
        lib.gg_mesh_bind_forward_ex.argtypes = [C.c_int32] * 3 + [vp] * 12 + [i32, vp]
        lib.gg_mesh_bind_backward_ex.argtypes = [C.c_int32] * 3 + [vp] * 17 + [i32, vp]
        lib.gg_cast_rays_workspace_bytes.argtypes = [C.c_int32, C.c_int32, P(sz), P(C.c_int64)]
        lib.gg_cast_rays_from_point.argtypes = [C.c_int32] * 3 + [vp] * 6 + [i64, C.c_int32, vp, vp, i32, vp]
        lib.gg_nvls_allreduce_f32.argtypes = [vp, vp, C.c_int32, C.c_int32, i64, i64, C.c_float, C.c_int32, C.c_int32, i32, vp]
        lib.gg_photometric_workspace_bytes.argtypes = [C.c_int32, C.c_int32, P(sz)]
        lib.gg_photometric_forward.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_int32, i32, vp]
        lib.gg_photometric_l1_u8.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_float, vp, vp, i32, vp]
        lib.gg_photometric_reduce.argtypes = [C.c_int32, C.c_int32, vp, C.c_float, vp, i32, vp]
        lib.gg_photometric_backward.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, i32, vp]
        lib.gg_kernel_timing.argtypes = [i32]
        lib.gg_kernel_count.restype = C.c_int
        lib.gg_kernel_name.restype = C.c_char_p
        lib.gg_kernel_name.argtypes = [i32]
        lib.gg_kernel_times.argtypes = [P(C.c_float)]
        for name in ("gg_forward_workspace_bytes", "gg_instance_workspace_bytes", "gg_backward_workspace_bytes",
                     "gg_forward_project", "gg_forward_color", "gg_forward_render", "gg_backward",
                     "gg_forward_overflow_check", "gg_forward_render_late_color", "gg_gate_signal", "gg_mark_visible", "gg_debug_read_geom", "gg_debug_read_binning", "gg_debug_lazy_phase_counters", "gg_kernel_timing", "gg_kernel_times",
                     "gg_mesh_bind_workspace_bytes", "gg_mesh_bind_forward", "gg_mesh_bind_backward",
                     "gg_mesh_bind_forward_ex", "gg_mesh_bind_backward_ex", "gg_cast_rays_workspace_bytes",
                     "gg_cast_rays_from_point", "gg_nvls_allreduce_f32",
                     "gg_photometric_workspace_bytes", "gg_photometric_forward", "gg_photometric_backward", "gg_photometric_reduce", "gg_photometric_l1_u8"):
            getattr(lib, name).restype = C.c_int
        if lib.gg_abi_version() != 1:
            raise RuntimeError("gaussian-garments_b200: libgg_raster.so ABI mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().gg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"gaussian-garments_b200: {what} failed (code {rc}): {msg}")


EXPORTED_SYMBOLS = [
    "gg_abi_version", "gg_version", "gg_last_error", "gg_launch_count", "gg_forward_workspace_bytes",
    "gg_instance_workspace_bytes", "gg_backward_workspace_bytes", "gg_forward_project", "gg_forward_color",
    "gg_forward_render", "gg_forward_render_late_color", "gg_gate_signal", "gg_forward_overflow_check", "gg_backward", "gg_mark_visible", "gg_debug_read_geom", "gg_debug_read_binning",
    "gg_debug_lazy_phase_counters", "gg_kernel_timing",
    "gg_kernel_count", "gg_kernel_name", "gg_kernel_times", "gg_mesh_bind_workspace_bytes", "gg_mesh_bind_forward",
    "gg_mesh_bind_backward", "gg_mesh_bind_forward_ex", "gg_mesh_bind_backward_ex", "gg_cast_rays_workspace_bytes",
    "gg_cast_rays_from_point", "gg_nvls_allreduce_f32", "gg_photometric_workspace_bytes", "gg_photometric_forward", "gg_photometric_backward",
    "gg_photometric_reduce", "gg_photometric_l1_u8",
]


def kernel_timing(enable: bool):
    check(load().gg_kernel_timing(1 if enable else 0), "gg_kernel_timing")


def kernel_times() -> dict:
    lib = load()
    n = lib.gg_kernel_count()
    buf = (C.c_float * n)()
    check(lib.gg_kernel_times(buf), "gg_kernel_times")
    return {lib.gg_kernel_name(i).decode(): float(buf[i]) for i in range(n) if buf[i] >= 0}


def launch_count(reset: bool = False) -> int:
    return int(load().gg_launch_count(1 if reset else 0))
