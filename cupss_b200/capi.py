"""ctypes binding of the flat C facade (tools/cupss_capi.h) over the C++ ``evolver`` API.

The facade mirrors /root/reference/inc/cupss/evolver.h:31-85 call for call, so a
python test reads like one of the reference's example ``main()``s.  The same
binding drives the product library (``lib/libcupss.so`` -> CUDA path) and, in
tests / bench / smoke only, the oracle builds of the reference under
``oracle/_ref`` (the caller passes the path; nothing here knows about oracles).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.path.join(REPO_ROOT, "lib", "libcupss.so")

RUN_CPU, RUN_GPU = 0, 1


class Pres(C.Structure):
    """``struct pres`` (/root/reference/inc/cupss/defines.h:31-39)."""

    _fields_ = [("preFactor", C.c_float), ("q2n", C.c_int), ("iqx", C.c_int),
                ("iqy", C.c_int), ("iqz", C.c_int), ("invq", C.c_int)]


_SIGS = {
    "cupss_capi_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]),
    "cupss_capi_destroy": (None, [C.c_void_p]),
    "cupss_capi_create_field": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "cupss_capi_add_parameter": (C.c_int, [C.c_void_p, C.c_char_p, C.c_float]),
    "cupss_capi_add_equation": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cupss_capi_add_noise": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "cupss_capi_create_term": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(Pres), C.c_int, C.POINTER(C.c_char_p), C.c_int]),
    "cupss_capi_create_from_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cupss_capi_prepare_problem": (None, [C.c_void_p]),
    "cupss_capi_advance_time": (C.c_int, [C.c_void_p, C.c_int]),
    "cupss_capi_copy_all_data_to_host": (None, [C.c_void_p]),
    "cupss_capi_write_out": (None, [C.c_void_p]),
    "cupss_capi_set_output_field": (None, [C.c_void_p, C.c_char_p, C.c_int]),
    "cupss_capi_update_parameter": (C.c_int, [C.c_void_p, C.c_char_p, C.c_float]),
    "cupss_capi_get_parameter": (C.c_float, [C.c_void_p, C.c_char_p]),
    "cupss_capi_get_timestep": (C.c_int, [C.c_void_p]),
    "cupss_capi_get_time": (C.c_float, [C.c_void_p]),
    "cupss_capi_set_write_precision": (None, [C.c_void_p, C.c_int]),
    "cupss_capi_field_real": (C.POINTER(C.c_float), [C.c_void_p, C.c_char_p]),
    "cupss_capi_field_comp": (C.POINTER(C.c_float), [C.c_void_p, C.c_char_p]),
    "cupss_capi_initialize_uniform": (None, [C.c_void_p, C.c_char_p, C.c_float]),
    "cupss_capi_initialize_droplet": (None, [C.c_void_p, C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]),
    "cupss_capi_add_droplet": (None, [C.c_void_p, C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]),
    "cupss_capi_initialize_half_system": (None, [C.c_void_p, C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_int]),
    "cupss_capi_initialize_from_file": (None, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char]),
    "cupss_capi_dump_plan": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "cupss_capi_set_mirror_callback": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "cupss_capi_print_information": (None, [C.c_void_p]),
    "cupss_capi_copy_host_to_device": (None, [C.c_void_p, C.c_char_p]),
    "cupss_capi_set_fourier_callback": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
}

# entry points that only the product build of the facade exports (tools/cupss_capi.h, CUPSS_B200_PRODUCT)
_PRODUCT_SIGS = {
    "cupss_capi_engine_plan": (C.c_void_p, [C.c_void_p]),
    "cupss_capi_set_noise_seed": (None, [C.c_void_p, C.c_ulonglong]),
    "cupss_capi_get_noise_seed": (C.c_ulonglong, [C.c_void_p]),
    "cupss_capi_set_partition": (None, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
}

# the engine's C ABI (include/cupss_b200.h): only the measurement / multi-GPU hooks are bound here,
# the rest is reached through the C++ evolver exactly like a user's main() would
ENGINE_LIB = os.path.join(REPO_ROOT, "cupss_b200", "lib", "libcupss_b200.so")
_ENGINE_SIGS = {
    "cupss_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]),
    "cupss_b200_destroy": (None, [C.c_void_p]),
    "cupss_b200_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "cupss_b200_set_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "cupss_b200_add_field": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "cupss_b200_set_implicit": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Pres), C.c_int]),
    "cupss_b200_clear_terms": (C.c_int, [C.c_void_p, C.c_int]),
    "cupss_b200_add_term": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Pres), C.c_int, C.POINTER(C.c_int), C.c_int]),
    "cupss_b200_set_noise": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Pres), C.c_ulonglong]),
    "cupss_b200_set_dealias_rule": (C.c_int, [C.c_void_p, C.c_int]),
    "cupss_b200_finalize": (C.c_int, [C.c_void_p]),
    "cupss_b200_upload_real": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cupss_b200_download_real": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cupss_b200_download_comp": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cupss_b200_step": (C.c_int, [C.c_void_p, C.c_int]),
    "cupss_b200_sync": (C.c_int, [C.c_void_p]),
    "cupss_b200_step_stage": (C.c_int, [C.c_void_p, C.c_int]),
    "cupss_b200_real_view_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cupss_b200_real_view_commit": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "cupss_b200_comp_view_begin": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "cupss_b200_comp_view_commit": (C.c_int, [C.c_void_p, C.c_int]),
    "cupss_b200_field_alias": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cupss_b200_time_steps": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "cupss_b200_profile_step": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "cupss_b200_launches_per_step": (C.c_int, [C.c_void_p]),
    "cupss_b200_bytes_per_step": (C.c_double, [C.c_void_p]),
    "cupss_b200_comm_bytes_per_step": (C.c_double, [C.c_void_p]),
    "cupss_b200_device_spectrum": (C.c_void_p, [C.c_void_p, C.c_int]),
    "cupss_b200_jit_selftest": (C.c_int, [C.c_char_p, C.c_int]),
    "cupss_b200_last_error": (C.c_char_p, []),
}

_LIBS: dict[str, C.CDLL] = {}
_ENGINE = None


def load_engine() -> C.CDLL:
    """dlopen the CUDA engine (C ABI).  Fails loudly if it has not been built: there is no fallback."""
    global _ENGINE
    if _ENGINE is None:
        if not os.path.exists(ENGINE_LIB):
            raise FileNotFoundError(f"{ENGINE_LIB} not found -- run __graft_entry__.build() first; the CUDA engine has no CPU fallback")
        lib = C.CDLL(ENGINE_LIB, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _ENGINE_SIGS.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export what include/cupss_b200.h declares
            fn.restype = res
            fn.argtypes = args
        _ENGINE = lib
    return _ENGINE


def engine_symbols():
    return sorted(_ENGINE_SIGS)


class EngineError(RuntimeError):
    pass


def engine_check(code: int, what: str = ""):
    if code != 0:
        raise EngineError(f"{what}: {load_engine().cupss_b200_last_error().decode()} (code {code})")


def load_facade(path: str) -> C.CDLL:
    """dlopen a library exporting the facade and attach prototypes.  Fails loudly if absent."""
    path = os.path.abspath(path)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in _PRODUCT_SIGS.items():
        fn = getattr(lib, name, None)   # absent from the oracle builds
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    _LIBS[path] = lib
    return lib


class Evolver:
    """Python spelling of the reference's ``evolver`` (same method names, camelCase kept)."""

    def __init__(self, device: int, sx: int, sy: int = 1, sz: int = 1, dx: float = 1.0, dy: float = 1.0,
                 dz: float = 1.0, dt: float = 0.1, write_every: int = 1 << 30, lib: str | None = None):
        self._lib = load_facade(lib or PRODUCT_LIB)
        self.sx, self.sy, self.sz = int(sx), int(sy), int(sz)
        self.n = self.sx * self.sy * self.sz
        self._h = self._lib.cupss_capi_create(int(device), sx, sy, sz, dx, dy, dz, dt, int(write_every))
        if not self._h:
            raise RuntimeError("evolver construction failed")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cupss_capi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- B200 engine access (product library only) ---------------------------
    def enginePlan(self):
        return self._lib.cupss_capi_engine_plan(self._h)

    def setNoiseSeed(self, seed: int):
        self._lib.cupss_capi_set_noise_seed(self._h, int(seed))

    def getNoiseSeed(self) -> int:
        return int(self._lib.cupss_capi_get_noise_seed(self._h))

    def setPartition(self, rank: int, nranks: int, nccl_id: bytes):
        self._lib.cupss_capi_set_partition(self._h, int(rank), int(nranks), nccl_id)

    def timeSteps(self, nsteps: int) -> float:
        """Run ``nsteps`` steps and return the CUDA-event time (ms) measured on the engine's own stream."""
        ms = C.c_float(0)
        engine_check(load_engine().cupss_b200_time_steps(self.enginePlan(), int(nsteps), C.byref(ms)), "time_steps")
        return float(ms.value)

    def profileStep(self):
        """One step with an event after every launch: [(name, ms, algorithmic_bytes)]."""
        n = C.c_int(0)
        names = C.create_string_buffer(64 * 64)
        ms = (C.c_float * 64)()
        by = (C.c_double * 64)()
        engine_check(load_engine().cupss_b200_profile_step(self.enginePlan(), 64, names, ms, by, C.byref(n)), "profile_step")
        return [(names.raw[64 * i:64 * i + 64].split(b"\0")[0].decode(), float(ms[i]), float(by[i])) for i in range(n.value)]

    def launchesPerStep(self) -> int:
        return load_engine().cupss_b200_launches_per_step(self.enginePlan())

    def bytesPerStep(self) -> float:
        return load_engine().cupss_b200_bytes_per_step(self.enginePlan())

    def commBytesPerStep(self) -> float:
        return load_engine().cupss_b200_comm_bytes_per_step(self.enginePlan())

    def sync(self):
        engine_check(load_engine().cupss_b200_sync(self.enginePlan()), "sync")

    # --- system declaration -------------------------------------------------
    def createField(self, name: str, dynamic: bool) -> int:
        return self._lib.cupss_capi_create_field(self._h, name.encode(), int(bool(dynamic)))

    def addParameter(self, name: str, value: float) -> int:
        return self._lib.cupss_capi_add_parameter(self._h, name.encode(), float(value))

    def addEquation(self, eq: str) -> int:
        return self._lib.cupss_capi_add_equation(self._h, eq.encode())

    def addNoise(self, field: str, expr: str) -> int:
        return self._lib.cupss_capi_add_noise(self._h, field.encode(), expr.encode())

    def createTerm(self, field: str, prefactors, product) -> int:
        arr = (Pres * len(prefactors))(*[Pres(*p) for p in prefactors])
        names = (C.c_char_p * max(1, len(product)))(*[p.encode() for p in product])
        return self._lib.cupss_capi_create_term(self._h, field.encode(), arr, len(prefactors), names, len(product))

    def createFromFile(self, path: str) -> int:
        return self._lib.cupss_capi_create_from_file(self._h, path.encode())

    # --- dynamics -------------------------------------------------------------
    def prepareProblem(self):
        self._lib.cupss_capi_prepare_problem(self._h)

    def advanceTime(self, nsteps: int = 1) -> int:
        return self._lib.cupss_capi_advance_time(self._h, int(nsteps))

    def copyAllDataToHost(self):
        self._lib.cupss_capi_copy_all_data_to_host(self._h)

    def writeOut(self):
        self._lib.cupss_capi_write_out(self._h)

    def setOutputField(self, name: str, on: bool):
        self._lib.cupss_capi_set_output_field(self._h, name.encode(), int(bool(on)))

    def updateParameter(self, name: str, value: float) -> int:
        return self._lib.cupss_capi_update_parameter(self._h, name.encode(), float(value))

    def getParameter(self, name: str) -> float:
        return self._lib.cupss_capi_get_parameter(self._h, name.encode())

    def getCurrentTimestep(self) -> int:
        return self._lib.cupss_capi_get_timestep(self._h)

    def getCurrentTime(self) -> float:
        return self._lib.cupss_capi_get_time(self._h)

    def setWritePrecision(self, digits: int):
        self._lib.cupss_capi_set_write_precision(self._h, int(digits))

    # --- host mirrors -----------------------------------------------------------
    def _view(self, ptr) -> np.ndarray:
        if not ptr:
            raise KeyError("no such field")
        a = np.ctypeslib.as_array(ptr, shape=(self.n, 2))
        return a.reshape(self.sz, self.sy, self.sx, 2)

    def fieldReal(self, name: str) -> np.ndarray:
        """Writable view of ``fieldsReal[name]`` as float32 [sz, sy, sx, 2] (value in [..., 0])."""
        return self._view(self._lib.cupss_capi_field_real(self._h, name.encode()))

    def fieldFourier(self, name: str) -> np.ndarray:
        return self._view(self._lib.cupss_capi_field_comp(self._h, name.encode()))

    def setReal(self, name: str, values: np.ndarray):
        v = self.fieldReal(name)
        v[..., 0] = np.asarray(values, dtype=np.float32).reshape(self.sz, self.sy, self.sx)
        v[..., 1] = 0.0

    def real(self, name: str) -> np.ndarray:
        return np.array(self.fieldReal(name)[..., 0], copy=True)

    def comp(self, name: str) -> np.ndarray:
        a = self.fieldFourier(name)
        return (a[..., 0] + 1j * a[..., 1]).astype(np.complex64)

    # --- initialisers -----------------------------------------------------------
    def initializeUniform(self, name: str, value: float):
        self._lib.cupss_capi_initialize_uniform(self._h, name.encode(), float(value))

    def initializeDroplet(self, name, v_out, v_in, radius, width, cx, cy, cz):
        self._lib.cupss_capi_initialize_droplet(self._h, name.encode(), v_out, v_in, radius, width, cx, cy, cz)

    def addDroplet(self, name, value, radius, width, cx, cy, cz):
        self._lib.cupss_capi_add_droplet(self._h, name.encode(), value, radius, width, cx, cy, cz)

    def initializeHalfSystem(self, name, v1, v2, width, direction):
        self._lib.cupss_capi_initialize_half_system(self._h, name.encode(), v1, v2, width, direction)

    def initializeFromFile(self, name, path, skiprows=1, delimiter=","):
        self._lib.cupss_capi_initialize_from_file(self._h, name.encode(), path.encode(), skiprows, delimiter.encode())

    def setMirrorCallback(self, name: str, odd=False):
        """Install one of the facade's built-in host callbacks on a field of a RUN_CPU evolver: False / True = even / odd
        mirror boundary condition, 2 = the non-symmetric strip clamp (tools/cupss_capi.cpp)."""
        if self._lib.cupss_capi_set_mirror_callback(self._h, name.encode(), int(odd)) != 0:
            raise KeyError(name)

    def printInformation(self):
        self._lib.cupss_capi_print_information(self._h)

    def copyHostToDevice(self, name: str):
        """field::copyHostToDevice: push the (edited) host real array of a field to the device mid-run."""
        self._lib.cupss_capi_copy_host_to_device(self._h, name.encode())

    def setFourierCallback(self, name: str, kind: int = 0, device_flavour: bool = False):
        """Install one of the facade's built-in Fourier-space callbacks (field::callbackFourier; tools/cupss_capi.cpp)."""
        rc = self._lib.cupss_capi_set_fourier_callback(self._h, name.encode(), int(kind), 1 if device_flavour else 0)
        if rc == 1:
            raise KeyError(name)
        if rc != 0:
            raise ValueError(f"Fourier callback kind {kind} (device_flavour={device_flavour}) is not available in this library")

    def dumpPlan(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        n = self._lib.cupss_capi_dump_plan(self._h, buf, len(buf))
        if n < 0:
            buf = C.create_string_buffer(-n)
            n = self._lib.cupss_capi_dump_plan(self._h, buf, len(buf))
        return buf.value.decode()
