"""Drop-in `data_set` module: same classes / attributes / file format as the reference's
src/data_set.py.  Minibatch selection stays bit-exact with the reference: the indices are
CPython's `random.choices(range(K), cum_weights=1..K, k=B)` stream, generated natively by
`epde_minibatch_indices` from `random.getstate()` (and the advanced state is put back with
`random.setstate`), or by `random.choices` itself when EPDE_PY_MINIBATCH=1.
"""
import ctypes as C
import os
import random

import numpy as np
import torch


def _native_choices(K, B):
    from eigenpde_nn_b200 import _lib
    version, state, gauss = random.getstate()
    mt = (C.c_uint32 * 624)(*state[:624])
    pos = C.c_int32(state[624])
    out = np.empty(B, dtype=np.int64)
    _lib.check(_lib.lib().epde_minibatch_indices(mt, C.byref(pos), K, B, out.ctypes.data_as(C.c_void_p)),
               "epde_minibatch_indices")
    random.setstate((version, tuple(mt) + (pos.value,), gauss))
    return out


class data_set():
    """Trajectory data + importance weights (reference: data_set.py:5-88)."""

    def __init__(self, xvec, weights):
        self.X_vec = torch.from_numpy(np.ascontiguousarray(xvec)).double()
        self.weights = torch.from_numpy(np.ascontiguousarray(weights)).double()
        self.K = self.X_vec.shape[0]
        self.tot_dim = self.X_vec.shape[1]
        self.active_index = range(self.K)
        self.active_indices = range(self.K)
        self.batch_size = self.K
        self.cum_weights = np.arange(1, self.K + 1)
        print('[Info]  Range of weights: [%.3e, %.3e]' % (self.weights.min(), self.weights.max()))

    @classmethod
    def from_file(cls, states_filename):
        """Header line `K d`, then K rows of d coordinates + 1 weight (data_generator.py:85)."""
        table = np.loadtxt(states_filename, skiprows=1, ndmin=2)
        print("[Info] %d sampled data loaded from file: %s" % (table.shape[0], states_filename))
        return cls(table[:, :-1], table[:, -1])

    def weights_minibatch(self):
        return self.weights[self.active_indices]

    def draw_indices(self, batch_size):
        """The reference's index stream (data_set.py:67), as an int64 array."""
        if os.environ.get("EPDE_PY_MINIBATCH") == "1":
            return np.asarray(random.choices(range(self.K), cum_weights=self.cum_weights, k=batch_size), dtype=np.int64)
        return _native_choices(self.K, batch_size)

    def generate_minibatch(self, batch_size, minibatch_flag=True):
        if minibatch_flag:
            self.active_indices = self.draw_indices(batch_size)
            self.batch_size = batch_size
        else:
            if batch_size != self.K:
                print("\tWarning: all states (%d) are selected. Value of batch_size (%d) is ingored." % (self.K, batch_size))
            self.active_indices = range(self.K)
            self.batch_size = self.K
        rows = self.X_vec if isinstance(self.active_indices, range) else self.X_vec[self.active_indices]
        self.x_batch = torch.reshape(rows, (self.batch_size, self.tot_dim)).clone()
        self.x_batch.requires_grad = True

    def pre_processing_layer(self):
        return self.x_batch


class MD_data_set(data_set):
    """Molecular data: states aligned to a reference configuration before entering the nets
    (reference: data_set.py:91-160).  `align` / `pre_processing_layer` keep the reference's minibatch semantics for
    the loaders and evaluation scripts; for training, `eigen_solver.run` evaluates the (parameter-free) layer once
    for the whole data set on the GPU (`engine.md_align` -> `epde_md_align`: aligned rows + the metric of the layer's
    Jacobian), SURVEY.md 8(f) N1."""

    def __init__(self, xvec, weights):
        super().__init__(xvec, weights)
        self.dim = 3
        self.num_atoms = int(self.tot_dim / self.dim)

    @classmethod
    def from_file(cls, states_filename):
        return super(MD_data_set, cls).from_file(states_filename)

    def load_ref_state(self):
        """./data/ref_state.txt: line 1 = number of reference atoms, line 2 = their indices, then xyz rows."""
        path = './data/ref_state.txt'
        with open(path) as fh:
            head = fh.readlines()
        self.ref_num_atoms = int(head[0].strip())
        self.ref_index = [int(tok) for tok in head[1].split()]
        coords = np.loadtxt(path, skiprows=2)
        self.ref_x = torch.from_numpy(coords).double().reshape((self.ref_num_atoms, self.dim))
        return self

    def align(self):
        """Kabsch alignment of every state in the minibatch onto the reference atoms."""
        conf = self.x_batch.reshape((-1, self.num_atoms, self.dim))
        sel = conf[:, self.ref_index, :]
        centre = sel.mean(dim=1, keepdim=True)
        cov = torch.matmul((sel - centre).transpose(1, 2), self.ref_x)          # [B,3,3]
        u, _, vh = torch.linalg.svd(cov)
        flip = torch.eye(3, dtype=torch.float64).repeat(conf.shape[0], 1, 1)
        flip[:, 2, 2] = torch.sign(torch.linalg.det(torch.matmul(u, vh))).detach()
        rot = u @ flip @ vh
        return torch.matmul(conf - centre, rot).reshape((-1, self.tot_dim))

    def pre_processing_layer(self):
        return self.align()
