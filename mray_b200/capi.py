"""ctypes mirror of include/mray_b200.h.

Python-side names follow the reference's accelerator interface
(``BaseAcceleratorI::CastRays / CastVisibilityRays``, Tracer/AcceleratorC.h; layouts of
Tracer/TracerTypes.h) so that parity tests read like calls into the reference. Device arrays are
``torch`` CUDA tensors (torch is plumbing for device memory and streams only); host arrays are
numpy. The library is loaded from ``mray_b200/lib/libmray_b200.so`` — if it is missing or no CUDA
device exists every compute call raises :class:`MrbError`; nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmray_b200.so")

MRB_MEM_HOST, MRB_MEM_DEVICE = 0, 1
MRB_TRACE_WIDE, MRB_TRACE_BINARY_EXACT = 0, 1
MRB_TRACE_FRESH_OUTPUTS = 0x100   # OR-ed flag: outputs need not be read, misses get INVALID keys / zero hits / set bits
MRB_BUILD_DEFAULT, MRB_BUILD_REFERENCE_DELTA, MRB_BUILD_BINARY_ONLY, MRB_BUILD_SERIAL_COLLAPSE = 0, 1, 2, 4
INVALID_KEY = 0xFFFFFFFF

STATUS = {0: "MRB_OK", -1: "MRB_ERR_NO_DEVICE", -2: "MRB_ERR_INVALID_ARG", -3: "MRB_ERR_CUDA",
          -4: "MRB_ERR_OUT_OF_MEMORY", -5: "MRB_ERR_UNSUPPORTED"}

# numpy dtypes of the reference's device structs (Tracer/TracerTypes.h:L201-209,L276-283)
RAY_DTYPE = np.dtype([("pos", np.float32, 3), ("tMin", np.float32), ("dir", np.float32, 3), ("tMax", np.float32)])
HITKEY_DTYPE = np.dtype([("primKey", np.uint32), ("lightOrMatKey", np.uint32), ("transKey", np.uint32),
                         ("accelKey", np.uint32)])


class MrbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class AccelDesc(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("vertexCount", C.c_uint32),
                ("indices", C.c_void_p), ("triangleCount", C.c_uint32),
                ("memspace", C.c_int), ("primGroupId", C.c_uint32),
                ("rangeCount", C.c_uint32), ("primRanges", C.c_void_p),
                ("lightOrMatKeys", C.c_void_p), ("cullBackface", C.c_void_p),
                ("flags", C.c_uint32), ("vertexUVs", C.c_void_p), ("alphaTextureCount", C.c_uint32),
                ("alphaTextures", C.c_void_p), ("rangeAlphaMap", C.c_void_p)]


class AccelInfo(C.Structure):
    _fields_ = [("leafCount", C.c_uint32), ("nodeCount", C.c_uint32), ("wideNodeCount", C.c_uint32),
                ("duplicateCodes", C.c_uint32), ("aabb", C.c_float * 6), ("buildMs", C.c_float),
                ("deviceBytes", C.c_size_t)]


class InstanceDesc(C.Structure):
    _fields_ = [("accel", C.c_void_p), ("transform", C.c_float * 12), ("invTransform", C.c_float * 12),
                ("isIdentity", C.c_uint32), ("transformKey", C.c_uint32), ("accelKey", C.c_uint32),
                ("lightOrMatKeys", C.c_void_p)]


class SpectrumDesc(C.Structure):
    _fields_ = [("lut", C.c_void_p), ("lutResolution", C.c_uint32), ("observerXYZ", C.c_void_p),
                ("illuminantSPD", C.c_void_p), ("xyzToRGB", C.c_float * 9), ("wavelengthSampleMode", C.c_uint32)]


class RenderDesc(C.Structure):
    _fields_ = [("accel", C.c_void_p), ("vertexCount", C.c_uint32), ("triangleCount", C.c_uint32),
                ("vertexNormals", C.c_void_p), ("materialCount", C.c_uint32), ("albedo", C.c_void_p),
                ("lightCount", C.c_uint32), ("lightRadiance", C.c_void_p), ("lightTwoSided", C.c_void_p),
                ("camPosition", C.c_float * 3), ("camGaze", C.c_float * 3), ("camUp", C.c_float * 3),
                ("fovXY", C.c_float * 2), ("nearFar", C.c_float * 2),
                ("width", C.c_uint32), ("height", C.c_uint32), ("totalSPP", C.c_uint32), ("sampleMode", C.c_uint32),
                ("rrRange", C.c_uint32 * 2), ("filmFilterRadius", C.c_float), ("seed", C.c_uint64),
                ("maxPathCount", C.c_uint32), ("partitionRays", C.c_uint32),
                ("scene", C.c_void_p), ("instanceVertexNormals", C.POINTER(C.c_void_p)), ("spectrum", C.c_void_p),
                ("samplerType", C.c_uint32), ("sobolMatrices", C.c_void_p),
                ("textureCount", C.c_uint32), ("textures", C.c_void_p), ("albedoTexture", C.c_void_p),
                ("vertexUVs", C.c_void_p), ("instanceVertexUVs", C.POINTER(C.c_void_p)),
                ("fullResolution", C.c_uint32 * 2), ("regionMin", C.c_uint32 * 2), ("materialType", C.c_void_p),
                ("filmFilterType", C.c_uint32), ("sampleOffset", C.c_uint32), ("jobSPP", C.c_uint32), ("materialParams", C.c_void_p),
                ("vertexTBN", C.c_void_p), ("instanceVertexTBN", C.POINTER(C.c_void_p)),
                ("boundaryType", C.c_uint32), ("boundaryTexture", C.c_int32), ("boundaryRadiance", C.c_float * 3),
                ("boundaryTransform", C.c_void_p), ("sceneDiameter", C.c_float), ("luminanceRow", C.c_float * 3),
                ("normalTexture", C.c_void_p), ("textureLodMode", C.c_uint32)]


class SpectraLutDesc(C.Structure):
    _fields_ = [("cieXYZ", C.c_void_p), ("illuminantSPD", C.c_void_p), ("illuminantNormFactor", C.c_float),
                ("rgbToXYZ", C.c_float * 9), ("xyzToRGB", C.c_float * 9), ("resolution", C.c_uint32), ("optimizePassCount", C.c_uint32)]


class TextureDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint32),
                ("format", C.c_uint32), ("interp", C.c_uint32), ("edge", C.c_uint32), ("gamma", C.c_float), ("colorMatrix", C.c_void_p),
                ("mipCount", C.c_uint32), ("generateMips", C.c_uint32), ("mipFilterType", C.c_uint32), ("mipFilterRadius", C.c_float),
                ("clampResolution", C.c_uint32)]


TEX_INTERP = {"Nearest": 0, "Linear": 1}
TEX_EDGE = {"Wrap": 0, "Clamp": 1, "Mirror": 2}


def _fill_texture_desc(t, texture, keep):
    """texture = dict(data=[h, w, C] float32 / uint8 (level 0), interp=, edge=, gamma=, color_matrix=, mips=[level 1, level 2, ...]
    (explicit levels, each [max(h >> k, 1), max(w >> k, 1), C]), gen_mips=(filter name, radius) (TracerParameters.genMips + mipGenFilter))"""
    a = np.ascontiguousarray(texture["data"])
    if a.dtype != np.uint8:
        a = np.ascontiguousarray(a, np.float32)
    if a.ndim == 2:
        a = a[..., None]
    t.height, t.width, t.channels = a.shape
    t.format = 1 if a.dtype == np.uint8 else 0
    t.interp, t.edge = TEX_INTERP[texture.get("interp", "Linear")], TEX_EDGE[texture.get("edge", "Wrap")]
    t.gamma = float(texture.get("gamma", 1.0))
    t.mipCount = 1
    if texture.get("mips"):
        levels = [a.reshape(-1, t.channels)]
        for k, m in enumerate(texture["mips"]):
            m = np.ascontiguousarray(m, a.dtype)
            assert m.size == max(t.height >> (k + 1), 1) * max(t.width >> (k + 1), 1) * t.channels, "mip level has the wrong size"
            levels.append(m.reshape(-1, t.channels))
        a = np.ascontiguousarray(np.concatenate(levels, axis=0))
        t.mipCount = len(levels)
    if texture.get("gen_mips"):
        t.generateMips, t.mipFilterType, t.mipFilterRadius = 1, FILM_FILTERS[texture["gen_mips"][0]], float(texture["gen_mips"][1])
    if texture.get("clamp_res"):      # TracerParameters.clampedTexRes; filtered with gen_mips' filter (or clamp_filter, default Gaussian 2)
        t.clampResolution = int(texture["clamp_res"])
        if not texture.get("gen_mips") and texture.get("clamp_filter"):
            t.mipFilterType, t.mipFilterRadius = FILM_FILTERS[texture["clamp_filter"][0]], float(texture["clamp_filter"][1])
    keep.append(a)
    t.data = a.ctypes.data
    if texture.get("color_matrix") is not None:     # RGB -> RGB matrix into the global colour space (row-major 3x3)
        m = np.ascontiguousarray(texture["color_matrix"], np.float32).reshape(9)
        keep.append(m); t.colorMatrix = m.ctypes.data
    return a


def _texture_desc(texture):
    t, keep = TextureDesc(), []
    _fill_texture_desc(t, texture, keep)
    return t, keep


class RenderStats(C.Structure):
    _fields_ = [("pathsStarted", C.c_uint64), ("pathsCompleted", C.c_uint64), ("closestRays", C.c_uint64),
                ("shadowRays", C.c_uint64), ("iterations", C.c_uint64), ("neeSamples", C.c_uint64), ("finished", C.c_uint32)]


class KernelProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 5), ("samples", C.c_uint64 * 5)]


PROFILE_KINDS = ["trace_closest", "shade", "trace_any", "finish_reload", "trace_tail"]
SAMPLE_MODES = {"Pure": 0, "WithNextEventEstimation": 1, "WithNEEAndMIS": 2}
BOUNDARY_TYPES = {"Null": 0, "Skysphere_Spherical": 1, "Skysphere_CoOcta": 2}   # (L)Null / LightGroupSkysphere<...>
FILM_FILTERS = {"Box": 0, "Tent": 1, "Gaussian": 2, "Mitchell-Netravali": 3}   # FilterType::E (Core/TracerEnums.h:L162-173)
HOST_FN = C.CFUNCTYPE(None, C.c_void_p)
# the descriptor mirrors above are written for this ABI (include/mray_b200.h: MRB_ABI_VERSION)
MRB_ABI_VERSION = (0 << 16) | 8

# every symbol include/mray_b200.h declares (tests/test_capi_symbols.py checks the header against this)
_PROTOTYPES = {
    "mrb_abi_version": (C.c_uint32, []),
    "mrb_context_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mrb_context_destroy": (None, [C.c_void_p]),
    "mrb_context_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrb_context_synchronize": (C.c_int, [C.c_void_p]),
    "mrb_context_used_device_memory": (C.c_size_t, [C.c_void_p]),
    "mrb_context_total_device_memory": (C.c_size_t, [C.c_void_p]),
    "mrb_context_launch_count": (C.c_uint64, [C.c_void_p]),
    "mrb_context_last_fallback_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "mrb_last_error": (C.c_char_p, [C.c_void_p]),
    "mrb_accel_export_wide": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrb_context_set_alpha_seed": (C.c_int, [C.c_void_p, C.c_uint32]),
    "mrb_context_set_profiling": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32]),
    "mrb_context_get_profile": (C.c_int, [C.c_void_p, C.POINTER(KernelProfile)]),
    "mrb_accel_build": (C.c_int, [C.c_void_p, C.POINTER(AccelDesc), C.POINTER(C.c_void_p)]),
    "mrb_accel_destroy": (None, [C.c_void_p, C.c_void_p]),
    "mrb_accel_get_info": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(AccelInfo)]),
    "mrb_accel_export_lbvh": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 7),
    "mrb_cast_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "mrb_cast_visibility_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "mrb_scene_build": (C.c_int, [C.c_void_p, C.POINTER(InstanceDesc), C.c_uint32, C.POINTER(C.c_void_p)]),
    "mrb_scene_destroy": (None, [C.c_void_p, C.c_void_p]),
    "mrb_scene_export_tlas": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 6),
    "mrb_scene_cast_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "mrb_scene_cast_visibility_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "mrb_spectrum_create": (C.c_int, [C.c_void_p, C.POINTER(SpectrumDesc), C.POINTER(C.c_void_p)]),
    "mrb_spectrum_destroy": (None, [C.c_void_p, C.c_void_p]),
    "mrb_spectrum_sample_wavelengths": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]),
    "mrb_spectrum_convert_to_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]),
    "mrb_spectrum_upsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                                        C.c_int, C.c_int]),
    "mrb_sampler_generate": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p, C.c_int]),
    "mrb_renderer_create": (C.c_int, [C.c_void_p, C.POINTER(RenderDesc), C.POINTER(C.c_void_p)]),
    "mrb_renderer_destroy": (None, [C.c_void_p, C.c_void_p]),
    "mrb_renderer_iterate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "mrb_renderer_get_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]),
    "mrb_renderer_read_film": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mrb_renderer_film_device_ptr": (C.c_void_p, [C.c_void_p]),
    "mrb_renderer_set_spp_limit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "mrb_renderer_begin_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]),
    "mrb_renderer_run_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(RenderStats)]),
    "mrb_renderer_poll_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]),
    "mrb_renderer_film_handoff": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrb_renderer_reduce_peers": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32]),
    "mrb_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mrb_host_free": (None, [C.c_void_p, C.c_void_p]),
    "mrb_filter_sample": (C.c_int, [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p]),
    "mrb_texture_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "mrb_dist2d_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]),
    "mrb_dist2d_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]),
    "mrb_skysphere_convert": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]),
    "mrb_texture_luminance": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]),
    "mrb_texture_convert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrb_texture_chain_texels": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "mrb_texture_full_mip_count": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "mrb_texture_final_extent": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "mrb_texture_mip_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]),
    "mrb_texture_sample_lod": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "mrb_spectra_lut_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrb_multi_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                      C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mrb_binary_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]),
    "mrb_radix_sort_pairs_u64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "mrb_radix_sort_pairs_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
}

_lib = None


def build_library(verbose: bool = False) -> str:
    """Compiles mray_b200/csrc for sm_100a into mray_b200/lib/libmray_b200.so (nvcc cross-compiles
    without a GPU). Raises if nvcc fails."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libmray_b200.so failed:\n" + res.stdout)
    return LIB_PATH


def load_library():
    """Loads the CUDA extension. Fails loudly when it has not been built."""
    global _lib, LIB_PATH
    if _lib is None:
        if os.environ.get("MRB_LIB_PATH"):      # experiment builds (csrc/Makefile VARIANT=...)
            LIB_PATH = os.environ["MRB_LIB_PATH"]
        if not os.path.exists(LIB_PATH):
            raise MrbError(-1, f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        got = int(lib.mrb_abi_version())
        if got != MRB_ABI_VERSION:   # a stale prebuilt .so would be driven with mismatched descriptor layouts
            raise MrbError(-5, f"{LIB_PATH} has ABI {got:#x}, this module was written for {MRB_ABI_VERSION:#x}: rebuild it")
        _lib = lib
    return _lib


def _ptr(x):
    """Raw pointer of a numpy array, torch tensor or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


def _space(*xs):
    spaces = set()
    for x in xs:
        if x is None:
            continue
        spaces.add(MRB_MEM_HOST if isinstance(x, np.ndarray) else
                   (MRB_MEM_DEVICE if x.is_cuda else MRB_MEM_HOST))
    if len(spaces) != 1:
        raise ValueError("all arrays of one call must live in the same memory space")
    return spaces.pop()


class Context:
    """One device + stream + scratch arena (mrb_context)."""

    def __init__(self, device: int = 0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.mrb_context_create(device, C.byref(h))
        if st != 0:
            raise MrbError(st, (self.lib.mrb_last_error(None) or b"").decode())
        self.handle = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def check(self, st):
        if st != 0:
            raise MrbError(st, (self.lib.mrb_last_error(self.handle) or b"").decode())

    def set_stream(self, stream):
        """``stream``: a torch.cuda.Stream or a raw cudaStream_t integer (0 / None = default stream)."""
        raw = None if stream is None else (getattr(stream, "cuda_stream", stream) or None)
        self.check(self.lib.mrb_context_set_stream(self.handle, C.c_void_p(raw)))

    def synchronize(self):
        self.check(self.lib.mrb_context_synchronize(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mrb_context_launch_count(self.handle))

    @property
    def last_fallback_count(self) -> int:
        """Rays of the last wide cast that were re-traced by the exact binary fallback (syncs)."""
        return self.last_fallback_stats[0]

    @property
    def last_fallback_stats(self):
        """(uncertified total, near-tie, uncertified-leaf, full binary re-traversals) of the last wide cast."""
        out = (C.c_uint32 * 4)()
        self.check(self.lib.mrb_context_last_fallback_count(self.handle, out))
        return tuple(int(x) for x in out)

    def set_alpha_seed(self, seed: int):
        """Seed of the stochastic alpha test of the next casts (mrb_context_set_alpha_seed); advances by one per cast."""
        self.check(self.lib.mrb_context_set_alpha_seed(self.handle, seed & 0xFFFFFFFF))

    def set_profiling(self, enabled: bool, iteration_stride: int = 16):
        """Sampled CUDA-event timing of the renderer's kernels (mrb_context_set_profiling)."""
        self.check(self.lib.mrb_context_set_profiling(self.handle, 1 if enabled else 0, iteration_stride))

    def get_profile(self):
        """{kind: (accumulated ms, samples)} since profiling was switched on (synchronises)."""
        p = KernelProfile()
        self.check(self.lib.mrb_context_get_profile(self.handle, C.byref(p)))
        return {k: (float(p.ms[i]), int(p.samples[i])) for i, k in enumerate(PROFILE_KINDS)}

    @property
    def used_device_memory(self) -> int:
        return int(self.lib.mrb_context_used_device_memory(self.handle))

    def sampler_generate(self, sampler_type, matrices, seeds, width, height, sample_index, initial_max_spp, dim_start, requests, out):
        """RNGGroupSobol / RNGGroupZSobol::GenerateNumbers (Tracer/Random.cu:L1017-1042,L1293-1318); out: u32[sum(requests), w*h]."""
        from . import sampling
        space = _space(matrices, seeds, out)
        t = sampling.sampler_code(sampler_type)
        req = (C.c_uint32 * len(requests))(*[int(r) for r in requests])
        self.check(self.lib.mrb_sampler_generate(self.handle, t, _ptr(matrices), _ptr(seeds), width, height, sample_index, initial_max_spp,
                                                 dim_start, req, len(requests), _ptr(out), space))

    def radix_sort_pairs(self, keys, values, bit_begin=0, bit_end=None):
        """DeviceAlgorithms::RadixSort<true,K,u32> (Device/CUDA/AlgRadixSortCUDA.h:L60-116), in place."""
        space = _space(keys, values)
        itemsize = keys.itemsize if isinstance(keys, np.ndarray) else keys.element_size()
        n = keys.shape[0]
        bit_end = itemsize * 8 if bit_end is None else bit_end
        fn = self.lib.mrb_radix_sort_pairs_u64 if itemsize == 8 else self.lib.mrb_radix_sort_pairs_u32
        self.check(fn(self.handle, _ptr(keys), _ptr(values), n, bit_begin, bit_end, space))

    def multi_partition(self, keys, indices, data_bits, batch_bits, max_partitions, only_batches=False):
        """RayPartitioner::MultiPartition on HOST numpy arrays (sorted in place). Returns
        (count, offsets[count+1], partition_keys[count])."""
        n = keys.shape[0]
        cnt = np.zeros(1, np.uint32); ofs = np.zeros(max_partitions + 1, np.uint32); pk = np.zeros(max_partitions, np.uint32)
        self.check(self.lib.mrb_multi_partition(self.handle, _ptr(keys), _ptr(indices), n, (C.c_uint32 * 2)(*data_bits),
                                                (C.c_uint32 * 2)(*batch_bits), 1 if only_batches else 0, max_partitions,
                                                cnt.ctypes.data, ofs.ctypes.data, pk.ctypes.data, MRB_MEM_HOST))
        c = int(cnt[0])
        return c, ofs[:c + 1].copy(), pk[:c].copy()

    def binary_partition(self, indices, flags):
        """RayPartitioner::BinaryPartition on HOST arrays: (partitioned indices, left count)."""
        n = indices.shape[0]
        out = np.zeros(n, np.uint32); left = np.zeros(1, np.uint32)
        flags = np.ascontiguousarray(flags, np.uint8)
        self.check(self.lib.mrb_binary_partition(self.handle, out.ctypes.data, left.ctypes.data, _ptr(indices), _ptr(flags),
                                                 flags.shape[0], n, MRB_MEM_HOST))
        return out, int(left[0])

    def close(self):
        if self.handle:
            self.lib.mrb_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Accelerator:
    """One concrete triangle accelerator — AcceleratorGroupLBVH<PrimGroupTriangle> with a single
    identity-transform instance (Tracer/AcceleratorLBVH.hpp:L433-899)."""

    def __init__(self, ctx: Context, positions, indices, prim_ranges=None, light_or_mat_keys=None,
                 cull_backface=None, prim_group_id: int = 0, flags: int = MRB_BUILD_DEFAULT,
                 vertex_uvs=None, alpha_textures=None, range_alpha_map=None):
        """alpha maps (SurfaceParams.alphaMaps): alpha_textures = list of dict(data=[h, w] or [h, w, c] float32 / uint8, interp=,
        edge=) read through their first channel, range_alpha_map = per prim range -1 or a texture index, vertex_uvs [V, 2]."""
        self.ctx = ctx
        d = AccelDesc()
        d.positions = _ptr(positions)
        d.vertexCount = positions.shape[0]
        d.indices = _ptr(indices)
        d.triangleCount = indices.shape[0]
        d.memspace = _space(positions, indices)
        d.primGroupId = prim_group_id
        self._keep = []
        if prim_ranges is not None:
            pr = np.ascontiguousarray(prim_ranges, np.uint32).reshape(-1, 2)
            d.rangeCount = pr.shape[0]
            d.primRanges = pr.ctypes.data
            self._keep.append(pr)
        if light_or_mat_keys is not None:
            lm = np.ascontiguousarray(light_or_mat_keys, np.uint32)
            d.lightOrMatKeys = lm.ctypes.data
            self._keep.append(lm)
        if cull_backface is not None:
            cb = np.ascontiguousarray(cull_backface, np.uint8)
            d.cullBackface = cb.ctypes.data
            self._keep.append(cb)
        d.flags = flags
        if range_alpha_map is not None:
            ram = np.ascontiguousarray(range_alpha_map, np.int32)
            self._keep.append(ram)
            d.rangeAlphaMap = ram.ctypes.data
            if vertex_uvs is not None:
                uv = vertex_uvs if not isinstance(vertex_uvs, np.ndarray) else np.ascontiguousarray(vertex_uvs, np.float32)
                self._keep.append(uv)
                d.vertexUVs = _ptr(uv)
            tarr = (TextureDesc * max(1, len(alpha_textures or [])))()
            for k, t in enumerate(alpha_textures or []):
                _fill_texture_desc(tarr[k], t, self._keep)
            self._keep.append(tarr)
            d.alphaTextureCount, d.alphaTextures = len(alpha_textures or []), C.cast(tarr, C.c_void_p)
        h = C.c_void_p()
        ctx.check(ctx.lib.mrb_accel_build(ctx.handle, C.byref(d), C.byref(h)))
        self.handle = h
        info = AccelInfo()
        ctx.check(ctx.lib.mrb_accel_get_info(ctx.handle, self.handle, C.byref(info)))
        self.info = info

    @property
    def leaf_count(self):
        return int(self.info.leafCount)

    @property
    def node_count(self):
        return int(self.info.nodeCount)

    def export_lbvh(self):
        """Binary LBVH artefacts in the reference layout (host numpy arrays)."""
        n, nn = self.leaf_count, self.node_count
        out = dict(morton=np.zeros(n, np.uint64), sorted_morton=np.zeros(n, np.uint64),
                   sorted_idx=np.zeros(n, np.uint32), nodes=np.zeros((nn, 3), np.uint32),
                   leaf_parent=np.zeros(n, np.uint32), boxes=np.zeros((nn, 6), np.float32),
                   leaf_aabb=np.zeros((n, 6), np.float32))
        self.ctx.check(self.ctx.lib.mrb_accel_export_lbvh(
            self.ctx.handle, self.handle, out["morton"].ctypes.data, out["sorted_morton"].ctypes.data,
            out["sorted_idx"].ctypes.data, out["nodes"].ctypes.data, out["leaf_parent"].ctypes.data,
            out["boxes"].ctypes.data, out["leaf_aabb"].ctypes.data))
        out["accel_aabb"] = np.array(list(self.info.aabb), np.float32)
        return out

    def export_wide(self):
        """The 8-wide tree as the kernels read it: (nodes uint32 [W, 20], triangle records float32 [N, 12])."""
        nodes = np.zeros((int(self.info.wideNodeCount), 20), np.uint32)
        tris = np.zeros((self.leaf_count, 12), np.float32)
        self.ctx.check(self.ctx.lib.mrb_accel_export_wide(self.ctx.handle, self.handle, nodes.ctypes.data, tris.ctypes.data))
        return nodes, tris

    def cast_rays(self, hit_keys, meta_hits, rays, ray_indices=None, mode=MRB_TRACE_WIDE):
        """BaseAcceleratorLBVH::CastRays (Tracer/AcceleratorLBVH.cu:L760-896): closest hit; writes
        hit_keys / meta_hits / rays.tMax at hit rays only."""
        space = _space(hit_keys, meta_hits, rays, ray_indices)
        total = rays.shape[0]
        count = total if ray_indices is None else ray_indices.shape[0]
        self.ctx.check(self.ctx.lib.mrb_cast_rays(self.ctx.handle, self.handle, _ptr(hit_keys), _ptr(meta_hits),
                                                  _ptr(rays), _ptr(ray_indices), count, total, space, mode))

    def cast_visibility_rays(self, visible_bits, rays, ray_indices=None, mode=MRB_TRACE_WIDE):
        """BaseAcceleratorLBVH::CastVisibilityRays (AcceleratorLBVH.cu:L898-1035): clears bit r when
        ray r is occluded."""
        space = _space(visible_bits, rays, ray_indices)
        total = rays.shape[0]
        count = total if ray_indices is None else ray_indices.shape[0]
        self.ctx.check(self.ctx.lib.mrb_cast_visibility_rays(self.ctx.handle, self.handle, _ptr(visible_bits),
                                                             _ptr(rays), _ptr(ray_indices), count, total, space, mode))

    def close(self):
        if self.handle and self.ctx.handle:
            self.ctx.lib.mrb_accel_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def light_key(index: int) -> int:
    """LightOrMatKey of light `index` (flag bit 31 set, Tracer/TracerTypes.h:L171-179)."""
    return 0x80000000 | (index & 0x1FFFFF)


class Spectrum:
    """SpectrumContextJakob2019 behind the C-ABI (Tracer/SpectrumContext.h:L76-155). `data` =
    mray_b200.spectral.load(); arrays are host numpy or device torch tensors like everywhere else."""

    def __init__(self, ctx: Context, data, wavelength_sample_mode="HyperbolicPBRT"):
        from . import spectral
        self.ctx = ctx
        d = SpectrumDesc()
        self._keep = [np.ascontiguousarray(data["lut"], np.float32), np.ascontiguousarray(data["observer"], np.float32),
                      np.ascontiguousarray(data["illuminant"], np.float32)]
        d.lut, d.lutResolution = self._keep[0].ctypes.data, int(data["resolution"])
        d.observerXYZ, d.illuminantSPD = self._keep[1].ctypes.data, self._keep[2].ctypes.data
        d.xyzToRGB = (C.c_float * 9)(*np.asarray(data["xyz_to_rgb"], np.float32).ravel())
        d.wavelengthSampleMode = (spectral.WAVELENGTH_SAMPLE_MODES[wavelength_sample_mode]
                                  if isinstance(wavelength_sample_mode, str) else int(wavelength_sample_mode))
        h = C.c_void_p()
        ctx.check(ctx.lib.mrb_spectrum_create(ctx.handle, C.byref(d), C.byref(h)))
        self.handle = h

    def sample_wavelengths(self, waves, pdfs, random_numbers):
        space = _space(waves, pdfs, random_numbers)
        self.ctx.check(self.ctx.lib.mrb_spectrum_sample_wavelengths(self.ctx.handle, self.handle, _ptr(waves), _ptr(pdfs),
                                                                    _ptr(random_numbers), random_numbers.shape[0], space))

    def convert_to_rgb(self, values, waves, pdfs):
        space = _space(values, waves, pdfs)
        self.ctx.check(self.ctx.lib.mrb_spectrum_convert_to_rgb(self.ctx.handle, self.handle, _ptr(values), _ptr(waves), _ptr(pdfs),
                                                                values.shape[0], space))

    def upsample(self, out, rgb, waves, is_radiance=False):
        space = _space(out, rgb, waves)
        uniform = 1 if rgb.ndim == 1 else 0
        self.ctx.check(self.ctx.lib.mrb_spectrum_upsample(self.ctx.handle, self.handle, _ptr(out), _ptr(rgb), uniform, _ptr(waves),
                                                          waves.shape[0], 1 if is_radiance else 0, space))

    def close(self):
        if self.handle and self.ctx.handle:
            self.ctx.lib.mrb_spectrum_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """(R)PathTracerRGB behind the C-ABI: StartRender / DoRenderWork / film read-out
    (TracerDLL/PathTracerRenderer.cu:L930-1355)."""

    def __init__(self, ctx: Context, accel, vertex_count, triangle_count, albedo, light_radiance,
                 camera, width, height, total_spp, sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=0,
                 vertex_normals=None, light_two_sided=None, film_filter_radius=1.0, near_far=(0.01, 1000.0),
                 max_path_count=0, partition_rays=False, instance_vertex_normals=None, spectrum=None, sampler="Independent",
                 textures=None, albedo_texture=None, vertex_uvs=None, instance_vertex_uvs=None,
                 full_resolution=None, region_min=(0, 0), material_type=None, film_filter="Gaussian",
                 sample_offset=0, job_spp=0, material_params=None, vertex_tbn=None, boundary=None, normal_texture=None, texture_lod_mode=0):
        """`accel` is an Accelerator, or a Scene (two-level; vertex_count / triangle_count / vertex_normals
        are then ignored and instance_vertex_normals may hold one array or None per instance).
        textures: list of dict(data=[h, w, 3|4] float32 or uint8 array, interp="Linear"|"Nearest",
        edge="Wrap"|"Clamp"|"Mirror"); albedo_texture: per material, -1 or an index into `textures`;
        vertex_uvs [V, 2] (instance_vertex_uvs: one array or None per instance of a Scene).
        boundary: None = (L)Null, or dict(type="Skysphere_Spherical"|"Skysphere_CoOcta", radiance=(r, g, b) | texture=index into
        `textures`, transform=[3, 4] local -> world or None, scene_diameter=0 (the reference's XZ diagonal))."""
        self.ctx, self.accel, self.spectrum = ctx, accel, spectrum
        self.width, self.height = width, height
        d = RenderDesc()
        self._keep = []
        if spectrum is not None:
            d.spectrum = spectrum.handle   # (R)PathTracerSpectral
        from . import sampling
        d.samplerType = sampling.sampler_code(sampler)
        if d.samplerType & 0xFF:
            mats = np.ascontiguousarray(sampling.sobol_matrices(), np.uint32)
            self._keep.append(mats)
            d.sobolMatrices = mats.ctypes.data
        if isinstance(accel, Scene):
            d.scene = accel.handle
            if instance_vertex_normals is not None:
                ptrs = (C.c_void_p * accel.count)()
                for k, nrm in enumerate(instance_vertex_normals):
                    if nrm is not None:
                        a = np.ascontiguousarray(nrm, np.float32); self._keep.append(a); ptrs[k] = a.ctypes.data
                self._keep.append(ptrs)
                d.instanceVertexNormals = C.cast(ptrs, C.POINTER(C.c_void_p))
            vertex_normals = None
        else:
            d.accel = accel.handle
            d.vertexCount, d.triangleCount = vertex_count, triangle_count
            if vertex_tbn is not None:   # [V, 4] world -> tangent-space quaternions (w, x, y, z): Quaternion::BarySLerp shading frames
                tq = np.ascontiguousarray(vertex_tbn, np.float32).reshape(-1, 4)
                self._keep.append(tq)
                d.vertexTBN = tq.ctypes.data

        def host(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dt)
            self._keep.append(a)
            return a.ctypes.data
        alb = np.asarray(albedo, np.float32).reshape(-1, 3)
        rad = np.asarray(light_radiance, np.float32).reshape(-1, 3)
        d.vertexNormals = host(vertex_normals, np.float32)
        d.materialCount, d.albedo = alb.shape[0], host(alb, np.float32)
        d.lightCount, d.lightRadiance = rad.shape[0], host(rad, np.float32)
        d.lightTwoSided = host(light_two_sided, np.uint8)
        d.camPosition = (C.c_float * 3)(*camera["eye"])
        d.camGaze = (C.c_float * 3)(*camera["gaze"])
        d.camUp = (C.c_float * 3)(*camera["up"])
        fy = float(np.deg2rad(camera["fov_y_deg"]))
        fx = float(2 * np.arctan(np.tan(fy / 2) * width / height))
        d.fovXY = (C.c_float * 2)(fx, fy)
        d.nearFar = (C.c_float * 2)(*near_far)
        d.width, d.height, d.totalSPP = width, height, total_spp
        d.sampleMode = SAMPLE_MODES[sample_mode] if isinstance(sample_mode, str) else int(sample_mode)
        d.rrRange = (C.c_uint32 * 2)(*rr_range)
        d.filmFilterRadius = film_filter_radius
        d.filmFilterType = FILM_FILTERS[film_filter] if isinstance(film_filter, str) else int(film_filter)
        d.sampleOffset, d.jobSPP = sample_offset, job_spp   # samples [sample_offset, sample_offset + total_spp) of a job_spp job
        d.seed = seed
        d.maxPathCount = max_path_count
        d.partitionRays = 1 if partition_rays else 0
        d.materialType = host(material_type, np.uint8)   # per material: 0 (Mt)Lambert, 1 (Mt)Reflect, 2 (Mt)Refract, 3 (Mt)Unreal
        if material_params is not None:                  # [material, 8]: Refract cauchyFront / cauchyBack, Unreal roughness / specular / metallic
            mp = np.ascontiguousarray(material_params, np.float32).reshape(-1, 8)
            assert mp.shape[0] == alb.shape[0], "material_params needs one row of 8 floats per material"
            d.materialParams = host(mp, np.float32)
        if full_resolution is not None:   # width x height is a region of a larger image
            d.fullResolution = (C.c_uint32 * 2)(*full_resolution)
            d.regionMin = (C.c_uint32 * 2)(*region_min)
        if textures:
            tarr = (TextureDesc * len(textures))()
            for k, t in enumerate(textures):
                _fill_texture_desc(tarr[k], t, self._keep)
            self._keep.append(tarr)
            d.textureCount, d.textures = len(textures), C.cast(tarr, C.c_void_p)
            d.albedoTexture = host(albedo_texture, np.int32)
            d.normalTexture = host(normal_texture, np.int32)    # per material: -1 or a texture holding tangent-space normals
            d.textureLodMode = texture_lod_mode   # 0 = mip level from UV-space gradients (reference host backend), 1 = texel-space (tex2DGrad)
        if isinstance(accel, Scene):
            if instance_vertex_uvs is not None:
                uptrs = (C.c_void_p * accel.count)()
                for k, uv in enumerate(instance_vertex_uvs):
                    if uv is not None:
                        a = np.ascontiguousarray(uv, np.float32); self._keep.append(a); uptrs[k] = a.ctypes.data
                self._keep.append(uptrs)
                d.instanceVertexUVs = C.cast(uptrs, C.POINTER(C.c_void_p))
        else:
            d.vertexUVs = host(vertex_uvs, np.float32)
        d.boundaryTexture = -1
        if boundary is not None:
            d.boundaryType = BOUNDARY_TYPES[boundary["type"]]
            d.boundaryTexture = int(boundary.get("texture", -1))
            d.boundaryRadiance = (C.c_float * 3)(*boundary.get("radiance", (0.0, 0.0, 0.0)))
            d.boundaryTransform = host(boundary.get("transform"), np.float32)
            d.sceneDiameter = float(boundary.get("scene_diameter", 0.0))
            if "luminance_row" in boundary:
                d.luminanceRow = (C.c_float * 3)(*boundary["luminance_row"])
        h = C.c_void_p()
        ctx.check(ctx.lib.mrb_renderer_create(ctx.handle, C.byref(d), C.byref(h)))
        self.handle = h

    def iterate(self, iterations=1):
        """DoRenderWork x iterations (asynchronous)."""
        self.ctx.check(self.ctx.lib.mrb_renderer_iterate(self.ctx.handle, self.handle, iterations))

    def set_spp_limit(self, spp_limit: int):
        """Latency mode: start camera paths only up to spp_limit samples per pixel (see mrb_renderer_set_spp_limit)."""
        self.ctx.check(self.ctx.lib.mrb_renderer_set_spp_limit(self.ctx.handle, self.handle, spp_limit))

    def stats(self) -> RenderStats:
        st = RenderStats()
        self.ctx.check(self.ctx.lib.mrb_renderer_get_stats(self.ctx.handle, self.handle, C.byref(st)))
        return st

    def begin_pass(self, region_min, region_size, sample_start, sample_count):
        """Samples [sample_start, sample_start + sample_count) of every pixel of the region (see mrb_renderer_begin_pass)."""
        rm = (C.c_uint32 * 2)(*region_min); rs = (C.c_uint32 * 2)(*region_size)
        self.ctx.check(self.ctx.lib.mrb_renderer_begin_pass(self.ctx.handle, self.handle, rm, rs, sample_start, sample_count))
        self.width, self.height = int(region_size[0]), int(region_size[1])

    def run_pass(self, chunk=4) -> RenderStats:
        """DoRenderWork until the current pass has finished (pipelined polling, no stream drain between polls)."""
        st = RenderStats()
        self.ctx.check(self.ctx.lib.mrb_renderer_run_pass(self.ctx.handle, self.handle, chunk, C.byref(st)))
        return st

    def poll_stats(self) -> RenderStats:
        """Non-blocking: the latest counter snapshot that has landed on the host (see mrb_renderer_poll_stats)."""
        st = RenderStats()
        self.ctx.check(self.ctx.lib.mrb_renderer_poll_stats(self.ctx.handle, self.handle, C.byref(st)))
        return st

    def film_handoff(self, host_dst, on_complete=None):
        """Asynchronous film hand-off into `host_dst` (a pinned numpy / torch host array of 4 x h x w floats)."""
        cb = HOST_FN(on_complete) if on_complete is not None else None
        if cb is not None:
            self._keep.append(cb)
        self.ctx.check(self.ctx.lib.mrb_renderer_film_handoff(self.ctx.handle, self.handle, _ptr(host_dst),
                                                              C.cast(cb, C.c_void_p) if cb is not None else None, None))

    def reduce_peers(self, peers):
        """Adds (and clears) the films of `peers` — Renderers of other devices — into this one's, over peer memory."""
        n = len(peers)
        pc = (C.c_void_p * n)(*[p.ctx.handle for p in peers]); pr = (C.c_void_p * n)(*[p.handle for p in peers])
        self.ctx.check(self.ctx.lib.mrb_renderer_reduce_peers(self.ctx.handle, self.handle, pc, pr, n))

    def read_film(self, clear=False):
        """(rgb_sum[h,w,3], weight[h,w]) host arrays, row 0 = bottom (RenderImageSection planes)."""
        out = np.zeros((4, self.height, self.width), np.float32)
        self.ctx.check(self.ctx.lib.mrb_renderer_read_film(self.ctx.handle, self.handle, out.ctypes.data, MRB_MEM_HOST,
                                                           1 if clear else 0))
        return np.ascontiguousarray(np.moveaxis(out[:3], 0, -1)), out[3]

    def film_device_ptr(self) -> int:
        return int(self.ctx.lib.mrb_renderer_film_device_ptr(self.handle))

    def render(self, batch=8, max_iterations=1_000_000):
        """Runs DoRenderWork until totalSPP*pixels paths have completed; returns the resolved image
        (sum radiance / sum weight, like MRay/RunCommand.cpp:L293-345) and the final stats."""
        st = self.run_pass(batch)
        rgb, w = self.read_film()
        return rgb / np.maximum(w, 1e-20)[..., None], st

    def close(self):
        if self.handle and self.ctx.handle:
            self.ctx.lib.mrb_renderer_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Two-level scene = BaseAcceleratorLBVH over accelerator instances (Tracer/AcceleratorLBVH.cu:L537-1035).
    instances: list of (Accelerator, transform 3x4 or None for identity[, light_or_mat_keys]). accelKey of
    instance i = i, transformKey = i + 1 (0 for identity). The optional third entry gives the instance its own
    LightOrMatKey per prim range (instances of one accelerator with different materials)."""

    def __init__(self, ctx: Context, instances):
        self.ctx = ctx
        self._keep = [t[0] for t in instances]
        arr = (InstanceDesc * len(instances))()
        self.transforms = []
        for i, t in enumerate(instances):
            acc, m = t[0], t[1]
            d = arr[i]
            d.accel = acc.handle
            if len(t) > 2 and t[2] is not None:
                k = np.ascontiguousarray(t[2], np.uint32)
                self._keep.append(k)
                d.lightOrMatKeys = k.ctypes.data
            ident = m is None
            m34 = np.eye(4, dtype=np.float64)[:3] if ident else np.asarray(m, np.float64).reshape(3, 4)
            inv = np.linalg.inv(np.vstack([m34, [0, 0, 0, 1]]))[:3]
            m32, i32 = m34.astype(np.float32), inv.astype(np.float32)
            d.transform = (C.c_float * 12)(*m32.ravel()); d.invTransform = (C.c_float * 12)(*i32.ravel())
            d.isIdentity = 1 if ident else 0
            d.transformKey = 0 if ident else i + 1
            d.accelKey = i
            self.transforms.append((m32, i32, ident))
        h = C.c_void_p()
        ctx.check(ctx.lib.mrb_scene_build(ctx.handle, arr, len(instances), C.byref(h)))
        self.handle = h
        self.count = len(instances)

    def export_tlas(self):
        n, nn = self.count, max(1, self.count - 1)
        out = dict(instance_aabb=np.zeros((n, 6), np.float32), scene_aabb=np.zeros(6, np.float32), morton=np.zeros(n, np.uint64),
                   sorted_idx=np.zeros(n, np.uint32), nodes=np.zeros((nn, 3), np.uint32), boxes=np.zeros((nn, 6), np.float32))
        self.ctx.check(self.ctx.lib.mrb_scene_export_tlas(self.ctx.handle, self.handle, out["instance_aabb"].ctypes.data,
                                                          out["scene_aabb"].ctypes.data, out["morton"].ctypes.data,
                                                          out["sorted_idx"].ctypes.data, out["nodes"].ctypes.data, out["boxes"].ctypes.data))
        return out

    def cast_rays(self, hit_keys, meta_hits, rays, ray_indices=None, mode=MRB_TRACE_WIDE):
        space = _space(hit_keys, meta_hits, rays, ray_indices)
        total = rays.shape[0]
        count = total if ray_indices is None else ray_indices.shape[0]
        self.ctx.check(self.ctx.lib.mrb_scene_cast_rays(self.ctx.handle, self.handle, _ptr(hit_keys), _ptr(meta_hits), _ptr(rays),
                                                        _ptr(ray_indices), count, total, space, mode))

    def cast_visibility_rays(self, visible_bits, rays, ray_indices=None, mode=MRB_TRACE_WIDE):
        space = _space(visible_bits, rays, ray_indices)
        total = rays.shape[0]
        count = total if ray_indices is None else ray_indices.shape[0]
        self.ctx.check(self.ctx.lib.mrb_scene_cast_visibility_rays(self.ctx.handle, self.handle, _ptr(visible_bits), _ptr(rays),
                                                                   _ptr(ray_indices), count, total, space, mode))

    def close(self):
        if self.handle and self.ctx.handle:
            self.ctx.lib.mrb_scene_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def texture_sample(ctx: Context, texture, uv):
    """mrb_texture_sample: texture = dict(data=[h, w, 3|4] float32 / uint8, interp=, edge=), uv[n, 2] -> rgb[n, 3]."""
    t, keep = _texture_desc(texture)
    uv = np.ascontiguousarray(uv, np.float32)
    out = np.zeros((uv.shape[0], 3), np.float32)
    ctx.check(ctx.lib.mrb_texture_sample(ctx.handle, C.byref(t), uv.ctypes.data, uv.shape[0], out.ctypes.data))
    return out


def texture_final_extent(ctx: Context, texture):
    """mrb_texture_final_extent: (width, height, mip count) after clampResolution / generateMips."""
    t, keep = _texture_desc(texture)
    e = (C.c_uint32 * 3)()
    ctx.check(ctx.lib.mrb_texture_final_extent(C.byref(t), e))
    return int(e[0]), int(e[1]), int(e[2])


def texture_mip_chain(ctx: Context, texture):
    """mrb_texture_mip_chain: (chain [total texels, C] in the texture's dtype, mip count, (width, height)) — resolution clamp applied,
    kept levels colour converted, generated levels after them."""
    t, keep = _texture_desc(texture)
    w, h, count = texture_final_extent(ctx, texture)
    texels = ctx.lib.mrb_texture_chain_texels(w, h, count)
    out = np.zeros((texels, t.channels), keep[0].dtype)
    got = C.c_uint32(0)
    ctx.check(ctx.lib.mrb_texture_mip_chain(ctx.handle, C.byref(t), out.ctypes.data, C.byref(got)))
    assert got.value == count
    return out, count


def texture_sample_lod(ctx: Context, texture, uv, lod=None, dpdx=None, dpdy=None, lod_mode=0):
    """mrb_texture_sample_lod: uv[n, 2] with lod[n], or with gradients dpdx / dpdy [n, 2] -> rgb[n, 3]."""
    t, keep = _texture_desc(texture)
    uv = np.ascontiguousarray(uv, np.float32)
    out = np.zeros((uv.shape[0], 3), np.float32)
    lp = gp = None
    if lod is not None:
        la = np.ascontiguousarray(lod, np.float32); lp = la.ctypes.data
    else:
        ga = np.ascontiguousarray(np.concatenate([np.asarray(dpdx, np.float32), np.asarray(dpdy, np.float32)], axis=1)); gp = ga.ctypes.data
    ctx.check(ctx.lib.mrb_texture_sample_lod(ctx.handle, C.byref(t), uv.ctypes.data, C.c_void_p(lp), C.c_void_p(gp), lod_mode, uv.shape[0], out.ctypes.data))
    return out


def texture_convert(ctx: Context, texture):
    """mrb_texture_convert: level 0 after the upload-time gamma / colour-space conversion (and resolution clamp), same dtype."""
    t, keep = _texture_desc(texture)
    w, h, _ = texture_final_extent(ctx, texture)
    out = np.zeros((h, w, t.channels), keep[0].dtype)
    ctx.check(ctx.lib.mrb_texture_convert(ctx.handle, C.byref(t), out.ctypes.data))
    return out


ACES_CG_LUMINANCE_ROW = (float.fromhex("0x1.1614ep-2"), float.fromhex("0x1.58e6fep-1"), float.fromhex("0x1.d946e6p-5"))


def texture_luminance(ctx: Context, texture, luminance_row=ACES_CG_LUMINANCE_ROW):
    """mrb_texture_luminance (KCExtractLuminance): -> [h, w] float32."""
    t, keep = _texture_desc(texture)
    out = np.zeros((t.height, t.width), np.float32)
    ctx.check(ctx.lib.mrb_texture_luminance(ctx.handle, C.byref(t), (C.c_float * 3)(*luminance_row), out.ctypes.data))
    return out


def dist2d_build(ctx: Context, function):
    """mrb_dist2d_build (DistributionGroupPwC2D::Construct): function [h, w] -> (cdf_x [h, w], cdf_y [h])."""
    f = np.ascontiguousarray(function, np.float32)
    h, w = f.shape
    cx, cy = np.zeros((h, w), np.float32), np.zeros(h, np.float32)
    ctx.check(ctx.lib.mrb_dist2d_build(ctx.handle, f.ctypes.data, w, h, cx.ctypes.data, cy.ctypes.data, MRB_MEM_HOST))
    return cx, cy


def dist2d_sample(ctx: Context, cdf_x, cdf_y, xi):
    """mrb_dist2d_sample: xi [n, 2] -> [n, 4] = (u, v, SampleUV pdf, PdfUV(u, v))."""
    cx, cy = np.ascontiguousarray(cdf_x, np.float32), np.ascontiguousarray(cdf_y, np.float32)
    xi = np.ascontiguousarray(xi, np.float32).reshape(-1, 2)
    out = np.zeros((xi.shape[0], 4), np.float32)
    ctx.check(ctx.lib.mrb_dist2d_sample(ctx.handle, cx.ctypes.data, cy.ctypes.data, cx.shape[1], cx.shape[0], xi.ctypes.data, xi.shape[0],
                                        out.ctypes.data, MRB_MEM_HOST))
    return out


def skysphere_convert(ctx: Context, converter, dirs):
    """mrb_skysphere_convert: unit Y-up dirs [n, 3] -> [n, 8] (see the header)."""
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((dirs.shape[0], 8), np.float32)
    c = BOUNDARY_TYPES[converter] if isinstance(converter, str) else int(converter)
    ctx.check(ctx.lib.mrb_skysphere_convert(ctx.handle, c, dirs.ctypes.data, dirs.shape[0], out.ctypes.data, MRB_MEM_HOST))
    return out


def spectra_lut_generate(ctx: Context, cie_xyz, illuminant_spd, illuminant_norm, rgb_to_xyz, xyz_to_rgb, resolution=64, passes=15):
    """mrb_spectra_lut_generate (the reference's SpectraLUTGen on the device): -> (lut float32 [9 * res^3], whitepoint float64 [3])."""
    cie = np.ascontiguousarray(cie_xyz, np.float32).reshape(-1, 3); spd = np.ascontiguousarray(illuminant_spd, np.float32)
    assert cie.shape[0] == 471 and spd.shape[0] == 471
    d = SpectraLutDesc()
    d.cieXYZ, d.illuminantSPD, d.illuminantNormFactor = cie.ctypes.data, spd.ctypes.data, float(illuminant_norm)
    d.rgbToXYZ = (C.c_float * 9)(*np.asarray(rgb_to_xyz, np.float32).ravel()); d.xyzToRGB = (C.c_float * 9)(*np.asarray(xyz_to_rgb, np.float32).ravel())
    d.resolution, d.optimizePassCount = resolution, passes
    lut = np.zeros(9 * resolution ** 3, np.float32); wp = np.zeros(3, np.float64)
    ctx.check(ctx.lib.mrb_spectra_lut_generate(ctx.handle, C.byref(d), lut.ctypes.data, wp.ctypes.data))
    return lut, wp


def filter_sample(ctx: Context, film_filter, radius, xi):
    """The film filter on its own (mrb_filter_sample): xi [n, 2] -> [n, 4] = (offset x, offset y, pdf, Evaluate(offset))."""
    xi = np.ascontiguousarray(xi, np.float32).reshape(-1, 2)
    out = np.zeros((xi.shape[0], 4), np.float32)
    t = FILM_FILTERS[film_filter] if isinstance(film_filter, str) else int(film_filter)
    ctx.check(ctx.lib.mrb_filter_sample(ctx.handle, t, float(radius), xi.ctypes.data, xi.shape[0], out.ctypes.data))
    return out
