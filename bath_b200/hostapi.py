"""ctypes binding of libbathhost.so (include/bathhost.h): the host side of the path."""
import ctypes as C
import os

import numpy as np

from .build import host_library_path

OK = 0
EXPORTS = [
    "bathhost_model_read", "bathhost_model_count", "bathhost_model_destroy", "bathhost_model_get_info",
    "bathhost_model_nrows", "bathhost_model_rfv", "bathhost_model_tfv", "bathhost_model_codons",
    "bathhost_model_indel_pos", "bathhost_model_mat", "bathhost_model_consensus", "bathhost_length_model",
]


class ModelInfo(C.Structure):
    _fields_ = [("M", C.c_int32), ("max_length", C.c_int32), ("codon_table", C.c_int32), ("fsprob", C.c_float),
                ("evparam", C.c_float * 8), ("has_fs3_stats", C.c_int32), ("has_fs5_stats", C.c_int32),
                ("name", C.c_char * 128), ("acc", C.c_char * 64)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = host_library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m bath_b200.build`")
    L = C.CDLL(path)
    vp, fp, u8p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    L.bathhost_model_read.restype = C.c_int
    L.bathhost_model_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.bathhost_model_count.restype = C.c_int
    L.bathhost_model_count.argtypes = [C.c_char_p]
    L.bathhost_model_destroy.restype = None
    L.bathhost_model_destroy.argtypes = [vp]
    L.bathhost_model_get_info.restype = C.c_int
    L.bathhost_model_get_info.argtypes = [vp, C.POINTER(ModelInfo)]
    L.bathhost_model_nrows.restype = C.c_int
    L.bathhost_model_nrows.argtypes = [vp, C.c_int]
    for name, res in (("bathhost_model_rfv", fp), ("bathhost_model_tfv", fp), ("bathhost_model_codons", u8p),
                      ("bathhost_model_indel_pos", u8p)):
        f = getattr(L, name)
        f.restype, f.argtypes = res, [vp, C.c_int]
    L.bathhost_model_mat.restype = fp
    L.bathhost_model_mat.argtypes = [vp]
    L.bathhost_model_consensus.restype = C.c_char_p
    L.bathhost_model_consensus.argtypes = [vp]
    L.bathhost_length_model.restype = None
    L.bathhost_length_model.argtypes = [C.c_int, C.c_float, fp, fp]
    _lib = L
    return L


class QueryModel:
    """One query profile set up as bathsearch does (src/bathsearch.c:794-801): local mode,
    3- and 5-codon-length frameshift profiles in odds-ratio form."""

    def __init__(self, path, index=0, ct=0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.bathhost_model_read(os.fsencode(path), index, ct, C.byref(h))
        if st != OK:
            raise IOError(f"cannot read model {index} of {path}: status {st}")
        self.h = h
        info = ModelInfo()
        self.lib.bathhost_model_get_info(h, C.byref(info))
        self.M = info.M
        self.max_length = info.max_length
        self.fsprob = info.fsprob
        self.codon_table = info.codon_table
        self.evparam = list(info.evparam)
        self.name = info.name.decode()

    def close(self):
        if getattr(self, "h", None):
            self.lib.bathhost_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nrows(self, which=3):
        return self.lib.bathhost_model_nrows(self.h, which)

    def rfv(self, which=3):
        return np.ctypeslib.as_array(self.lib.bathhost_model_rfv(self.h, which), shape=(self.nrows(which), self.M + 1))

    def tfv(self, which=3):
        return np.ctypeslib.as_array(self.lib.bathhost_model_tfv(self.h, which), shape=(8, self.M + 1))

    def codons(self, which=3):
        mc = self.nrows(which) - 29
        flat = np.ctypeslib.as_array(self.lib.bathhost_model_codons(self.h, which), shape=((self.M + 1) * (mc + 1),))
        return flat[: (self.M + 1) * mc].reshape(self.M + 1, mc)

    def indel_pos(self, which=3):
        mc = self.nrows(which) - 29
        flat = np.ctypeslib.as_array(self.lib.bathhost_model_indel_pos(self.h, which), shape=((self.M + 1) * (mc + 1),))
        return flat[: (self.M + 1) * mc].reshape(self.M + 1, mc)

    def mat(self):
        return np.ctypeslib.as_array(self.lib.bathhost_model_mat(self.h), shape=(self.M + 1, 20))


def length_model(L_amino, nj=1.0):
    pm, pl = C.c_float(), C.c_float()
    load().bathhost_length_model(int(L_amino), float(nj), C.byref(pm), C.byref(pl))
    return pm.value, pl.value
