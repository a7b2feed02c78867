"""ctypes wrapper of oracle/_build/libplanexec.so (the test-only CPU interpreter of the engine's
plan; TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libplanexec.so")

PLAN_INFER, PLAN_HUBERT, PLAN_PITCH, PLAN_MEL, PLAN_KNN, PLAN_FEATURE = range(6)


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return LIB


class PlanExec:
    def __init__(self, data_dir):
        build()
        L = ctypes.CDLL(LIB)
        L.pe_create.restype = ctypes.c_void_p
        L.pe_error.restype = ctypes.c_char_p
        L.pe_buf_name.restype = ctypes.c_char_p
        L.pe_get.restype = ctypes.c_long
        for f in (L.pe_destroy, L.pe_error, L.pe_load, L.pe_set_index, L.pe_set_params, L.pe_run,
                  L.pe_num_ops, L.pe_num_bufs, L.pe_buf_name, L.pe_get, L.pe_plan_dims, L.pe_set_chain, L.pe_chain_stats):
            f.argtypes = None
        self.L = L
        self.h = ctypes.c_void_p(L.pe_create(data_dir.encode()))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.pe_error(self.h).decode())

    def load(self, which, path):
        self._chk(self.L.pe_load(self.h, ctypes.c_int(which), path.encode()))

    def set_index(self, rows, rate):
        rows = np.ascontiguousarray(rows, np.float32)
        self._chk(self.L.pe_set_index(self.h, rows.ctypes.data_as(ctypes.c_void_p),
                                      ctypes.c_int(rows.shape[0]), ctypes.c_int(rows.shape[1]),
                                      ctypes.c_float(rate)))

    def set_params(self, seed=0, noise_mode=1, index_k=8):
        self.L.pe_set_params(self.h, ctypes.c_uint64(seed), ctypes.c_int(noise_mode),
                             ctypes.c_int(index_k))

    def run(self, kind, pcm, sf16k=0, pitch_shift=0, skip_head=0, return_length=0):
        pcm = np.ascontiguousarray(pcm, np.float32)
        self._chk(self.L.pe_run(self.h, ctypes.c_int(kind), pcm.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_int(pcm.shape[0]), ctypes.c_int(sf16k),
                                ctypes.c_int(pitch_shift), ctypes.c_int(skip_head),
                                ctypes.c_int(return_length)))

    def set_chain(self, grid_main=148, grid_side=32, side_max_m=8):
        """Plans built from now on group small same-lane ops into persistent chains (chain.h); `run` then executes
        every chain phase by phase with the ops of a phase in REVERSE order."""
        self.L.pe_set_chain(self.h, ctypes.c_int(grid_main), ctypes.c_int(grid_side), ctypes.c_int(side_max_m))

    def chain_stats(self):
        out = (ctypes.c_int * 4)()
        self.L.pe_chain_stats(self.h, out)
        return dict(n_chains=out[0], n_phases=out[1], n_chain_ops=out[2], widest_phase=out[3])

    def names(self):
        return [self.L.pe_buf_name(self.h, ctypes.c_int(i)).decode()
                for i in range(self.L.pe_num_bufs(self.h))]

    def get(self, name, dtype=np.float32, cap=1 << 24):
        buf = np.empty(cap, dtype=np.float32)
        n = self.L.pe_get(self.h, name.encode(), buf.ctypes.data_as(ctypes.c_void_p),
                          ctypes.c_long(cap))
        if n < 0:
            raise KeyError(name)
        return buf[:n].view(dtype).copy()

    def dims(self):
        out = (ctypes.c_int * 8)()
        self.L.pe_plan_dims(self.h, out)
        return dict(hubert_T=out[0], hubert_C=out[1], f0_T=out[2], audio_len=out[3], knn_q=out[4],
                    n_ops=out[5], n_lanes=out[6])

    def close(self):
        if self.h:
            self.L.pe_destroy(self.h)
            self.h = None
