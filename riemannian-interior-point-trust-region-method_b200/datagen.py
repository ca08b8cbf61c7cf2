"""Synthetic NonnegPCA instances by the reference's generator law (src/NonnegPCA/generator.py:9-65,
config_dataset.yaml:6-8: dim=50, snr=0.5, delta=0.7).  The reference draws from the global, unseeded
NumPy RNG; here every instance has its own `RandomState(seed)` (seed = instance id) so sweeps are
reproducible and shardable: rank r of a sweep generates exactly its own instances.

    Z  = sqrt(snr) v v' + N/sqrt(dim), diag(N) ~ 2/sqrt(dim) * randn      (:9-31, NOT symmetrised)
    x0 = |u / ||u|||, u ~ U(0,1)^dim                                       (:46-51)
    y0 = 1                                                                 (:63)
"""
import numpy as np


def nonnegpca_instance(dim=50, snr=0.5, delta=0.7, seed=0):
    rs = np.random.RandomState(seed)
    samplesize = int(np.floor(delta * dim))
    support = rs.choice(dim, samplesize, replace=False)
    v = np.zeros(dim)
    v[support] = 1 / np.sqrt(samplesize)
    Z = np.sqrt(snr) * np.outer(v, v)
    noise = rs.randn(dim, dim) / np.sqrt(dim)
    for ii in range(dim):
        noise[ii, ii] = rs.randn() * 2 / np.sqrt(dim)
    Z = Z + noise
    x0 = rs.rand(dim)
    x0 = np.abs(x0 / np.linalg.norm(x0))
    return Z, x0, np.ones(dim)


def nonnegpca_batch(first_seed, count, dim=50, snr=0.5, delta=0.7, out=None):
    """(Z [count, dim, dim], x0 [count, dim], y0 [count, dim]) for seeds first_seed .. first_seed+count-1.
    `out` may hold preallocated (e.g. pinned) arrays to fill."""
    if out is None:
        Z = np.empty((count, dim, dim))
        x0 = np.empty((count, dim))
        y0 = np.empty((count, dim))
    else:
        Z, x0, y0 = out
    for i in range(count):
        Z[i], x0[i], y0[i] = nonnegpca_instance(dim, snr, delta, first_seed + i)
    return Z, x0, y0


def more_initial_points(x0, seed, count):
    """`count` further strictly feasible starting points for one instance (the reference ships one,
    `initx_a`; generator.py:46-51 is the law): |u/||u|||, u ~ U(0,1)^dim."""
    rs = np.random.RandomState(seed)
    pts = rs.rand(count, x0.shape[0])
    return np.abs(pts / np.linalg.norm(pts, axis=1, keepdims=True))


def nonnegpca_sweep(first_instance, instances, points_per_instance, dim=50, out=None):
    """A sweep of `instances` problem instances x `points_per_instance` initial points (instance-major pair order):
    Z [instances, dim, dim], x0 / y0 [instances * points_per_instance, dim].  Point 0 of an instance is the
    generator's own x0 (seed = instance id); points k >= 1 follow the same law from seed 2e9 + 16 * id + k (k < 16)."""
    if not 1 <= points_per_instance <= 16:
        raise ValueError("1 <= points_per_instance <= 16")
    pairs = instances * points_per_instance
    if out is None:
        Z, x0, y0 = np.empty((instances, dim, dim)), np.empty((pairs, dim)), np.empty((pairs, dim))
    else:
        Z, x0, y0 = out
    for i in range(instances):
        inst = first_instance + i
        Z[i], x0[i * points_per_instance], _ = nonnegpca_instance(dim, seed=inst)
        for k in range(1, points_per_instance):
            x0[i * points_per_instance + k] = more_initial_points(x0[0], 2000000000 + 16 * inst + k, 1)[0]
    y0[:] = 1.0
    return Z, x0, y0


def stableid_constraint_values(A, conspec):
    """g_i(A) for the constraint rows [kind, row, col, a, b] (src/StableIdentification/coordinator.py:108-130)."""
    g = np.empty(len(conspec))
    for i, (kind, r, c, a, b) in enumerate(conspec):
        v = A[int(r), int(c)]
        g[i] = (-v + a) if kind == 0 else ((v - a) if kind == 1 else (-(v - a) ** 2 + b))
    return g


def stableid_more_initial_points(base_points, conspec, count, seed=0, scale=0.02, margin=1e-3):
    """`count` strictly feasible starting points (J, R, Q) for a StableIdentification instance: the reference ships 20 per
    instance (found with RALM, src/StableIdentification/generator.py); a sweep over more takes base point k % len(base)
    plus a small random skew / symmetric / symmetric perturbation, halved until R, Q stay positive definite and every
    constraint keeps g_i(A) <= -margin * |g_i(A_base)|.  Point k < len(base) is the base point itself."""
    rs = np.random.RandomState(seed)
    out = []
    nb = len(base_points)
    for k in range(count):
        J0, R0, Q0 = (np.asarray(a, dtype=np.float64) for a in base_points[k % nb])
        if k < nb:
            out.append([J0.copy(), R0.copy(), Q0.copy()])
            continue
        g0 = stableid_constraint_values((J0 - R0) @ Q0, conspec)
        d = J0.shape[0]
        EJ, ER, EQ = rs.randn(d, d), rs.randn(d, d), rs.randn(d, d)
        t = scale
        while True:
            J = J0 + t * 0.5 * (EJ - EJ.T)
            R = R0 + t * 0.5 * (ER + ER.T) * np.linalg.norm(R0) / d
            Q = Q0 + t * 0.5 * (EQ + EQ.T) * np.linalg.norm(Q0) / d
            ok = np.linalg.eigvalsh(R).min() > 0 and np.linalg.eigvalsh(Q).min() > 0
            if ok:
                g = stableid_constraint_values((J - R) @ Q, conspec)
                ok = bool(np.all(g <= -margin * np.abs(g0)))
            if ok:
                break
            t *= 0.5
        out.append([J, R, Q])
    return out


def nonnegpca_sweep_device(first_instance, instances, points_per_instance, dim=50, snr=0.5, delta=0.7, device=0, stream=None):
    """The same sweep drawn ON the device (csrc/datagen.cuh: Philox4x32-10 keyed by the instance id, the reference
    generator's law): returns torch CUDA tensors Z [instances, dim, dim], x0 / y0 [instances * points, dim].  A different
    random stream than `nonnegpca_sweep` (NumPy's MT19937 is sequential); tests/helpers.py restates it bit for bit."""
    import ctypes as C
    import torch
    from . import _lib
    lib = _lib.load_library()
    dev = torch.device("cuda", device)
    Z = torch.empty((instances, dim, dim), dtype=torch.float64, device=dev)
    x0 = torch.empty((instances * points_per_instance, dim), dtype=torch.float64, device=dev)
    y0 = torch.empty_like(x0)
    _lib.check(lib.riptrm_generate_nonnegpca(device, dim, int(first_instance), instances, points_per_instance, float(snr),
                                             float(delta), Z.data_ptr(), x0.data_ptr(), y0.data_ptr(),
                                             C.c_void_p(stream) if stream else None))
    return Z, x0, y0
