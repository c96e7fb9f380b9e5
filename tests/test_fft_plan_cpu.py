"""CPU check of the hand-written FFT passes' host side (csrc/fft2d.cu): the PRODUCT's plan constants (radices, strides, table offsets:
a host-only hook returns what the kernels are compiled with) and its twiddle / stage tables drive a numpy re-enactment of the kernels'
index algebra — Stockham stages with eight or sixteen points per thread, the in-place stages of the pipelined strided pass with its
digit-reversed output map, the split / merge step of the real transforms — which is compared with numpy's FFT; and the padded
shared-memory indices of the line kernels are checked for bank conflicts over every stage of every plan."""
import ctypes as C

import numpy as np
import pytest

PLANS = [(m, 3) for m in range(3, 11)] + [(7, 4), (8, 4)]


def _plan(pdo, log2n, loge):
    out = np.zeros(20, dtype=np.int32)
    assert pdo.lib().pdo_debug_fft_plan(log2n, loge, C.c_void_p(out.ctypes.data)) == 0
    ns = int(out[0])
    return dict(NS=ns, R0=int(out[1]), T=int(out[2]), E=int(out[3]), R=[int(v) for v in out[4:4 + ns]], s=[int(v) for v in out[8:8 + ns]],
                npts=[int(v) for v in out[12:12 + ns]], toff=[int(v) for v in out[16:16 + ns]])


def _tables(pdo, n, loge):
    buf = np.zeros(4 * n + 16, dtype=np.complex128)
    total = pdo.lib().pdo_debug_fft_tables(n, loge, C.c_void_p(buf.ctypes.data), buf.size)
    assert total >= n
    return buf[:n].copy(), buf[n:total].copy()


def _bfly(a, sgn):
    r = len(a)
    k = np.arange(r)
    return np.exp(sgn * 2j * np.pi * np.outer(k, k) / r) @ a


def _stockham(x, P, tw, sgn, pad=None, conflicts=None):
    """stage i: butterfly b = q + s p reads x[b + (N/R) r], writes y[q + s (R p + k)] = (sum_r x_r w_R^rk) w_N^(s p k)"""
    n, T, E = len(x), P["T"], P["E"]
    src = x.astype(complex)
    for i in range(P["NS"]):
        R, s, NP, off = P["R"][i], P["s"][i], P["npts"][i], P["toff"][i]
        dst = np.zeros(n, complex)
        for u in range(E // R):
            rd, wr = [[] for _ in range(R)], [[] for _ in range(R)]
            for t in range(T):
                b = t + T * u
                q, p = b % s, b // s
                y = _bfly(np.array([src[b + (n // R) * r] for r in range(R)]), sgn)
                for k in range(R):
                    w = 1.0
                    if i < P["NS"] - 1 and k > 0:
                        w = tw[off + (k - 1) * NP + p]
                        w = np.conj(w) if sgn > 0 else w
                    dst[q + s * (R * p + k)] = y[k] * w
                    rd[k].append(b + (n // R) * k)
                    wr[k].append(q + s * (R * p + k))
            if pad is not None:
                for lst in rd + wr:
                    for g in range(0, len(lst), 8):      # a quarter warp: eight 16-byte accesses, eight 16-byte bank groups
                        banks = [pad(v) % 8 for v in lst[g:g + 8]]
                        conflicts.append(max(banks.count(bk) for bk in set(banks)))
        src = dst
    return src


@pytest.mark.parametrize("log2n,loge", PLANS)
def test_stockham_plan_and_stage_tables_reproduce_the_dft(pdo, log2n, loge):
    n = 1 << log2n
    P = _plan(pdo, log2n, loge)
    assert P["E"] == 1 << loge and P["T"] * P["E"] == n and int(np.prod(P["R"])) == n
    flat, tw = _tables(pdo, n, loge)
    assert np.abs(flat - np.exp(-2j * np.pi * np.arange(n) / n)).max() < 1e-15       # numpy's own exp (argument rounding near 2 pi) is the less accurate side
    assert np.array_equal(flat[1:], np.conj(flat[:0:-1]))      # exact mirror symmetry of the octant construction
    rng = np.random.default_rng(n + loge)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ps = 4 if P["R0"] >= 16 else 3
    conflicts = []
    got = _stockham(x, P, tw, -1, pad=lambda p: p + (p >> ps), conflicts=conflicts)
    ref = np.fft.fft(x)
    assert np.abs(got - ref).max() < 1e-13 * np.abs(ref).max()
    assert max(conflicts) == 1, "a stage of the line kernels would hit a shared-memory bank conflict"
    back = _stockham(ref, P, tw, +1)
    assert np.abs(back - n * x).max() < 1e-12 * n


@pytest.mark.parametrize("log2n", [7, 8, 9])
def test_inplace_stages_and_output_map_of_the_pipelined_pass(pdo, log2n):
    """fft_cols_pipe_kernel: butterfly b = hi S + lo on slots hi R S + r S + lo, factor w_N^(lo P k) from the same stage tables
    (index lo), results in digit-reversed slots: slot -> sum_i digit_i * stride_i."""
    n = 1 << log2n
    P = _plan(pdo, log2n, 3)
    _, tw = _tables(pdo, n, 3)
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a = x.copy()
    for i in range(P["NS"]):
        R, S, off = P["R"][i], P["npts"][i], P["toff"][i]
        for u in range(8 // R):
            for t in range(P["T"]):
                b = t + P["T"] * u
                hi, lo = b // S, b % S
                base = hi * R * S + lo
                y = _bfly(np.array([a[base + r * S] for r in range(R)]), -1)
                for k in range(R):
                    w = tw[off + (k - 1) * S + lo] if (i < P["NS"] - 1 and k > 0) else 1.0
                    a[base + k * S] = y[k] * w
    out = np.zeros(n, complex)
    for pos in range(n):
        rem, f = pos, 0
        for i in range(P["NS"]):
            d, rem = divmod(rem, P["npts"][i])
            f += d * P["s"][i]
        out[f] = a[pos]
    ref = np.fft.fft(x)
    assert np.abs(out - ref).max() < 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("nx", [16, 64, 512])
def test_split_and_merge_steps_of_the_real_transforms(pdo, nx):
    """fft_r2c_kernel / fft_c2r_kernel: one complex transform of M = nx/2 points on z_p = x_2p + i x_2p+1, merged into (r2c) or
    split from (c2r) the nx/2 + 1 modes with the product's flat table; c2r ignores the imaginary parts of modes 0 and nx/2."""
    M = nx // 2
    flat, _ = _tables(pdo, nx, 3) if nx <= 1024 else (None, None)
    rng = np.random.default_rng(nx)
    x = rng.standard_normal(nx)
    Z = np.fft.fft(x[0::2] + 1j * x[1::2])
    X = np.zeros(M + 1, complex)
    X[0], X[M] = Z[0].real + Z[0].imag, Z[0].real - Z[0].imag
    for k in range(1, M // 2):
        A, B = Z[k], Z[M - k]
        Ev = 0.5 * (A + np.conj(B))
        Od = complex(0.5 * (A.imag + B.imag), -0.5 * (A.real - B.real))
        G = Od * flat[k]
        X[k], X[M - k] = Ev + G, np.conj(Ev - G)
    X[M // 2] = np.conj(Z[M // 2])
    ref = np.fft.rfft(x)
    assert np.abs(X - ref).max() < 1e-14 * np.abs(ref).max()
    Xs = ref.copy()
    Xs[0] += 0.3j
    Xs[M] -= 0.7j
    Zp = np.zeros(M, complex)
    Zp[0] = complex(Xs[0].real + Xs[M].real, Xs[0].real - Xs[M].real)
    for k in range(1, M // 2):
        A, B = Xs[k], Xs[M - k]
        Ze, D = A + np.conj(B), A - np.conj(B)
        G = D * np.conj(flat[k])
        Zp[k] = complex(Ze.real - G.imag, Ze.imag + G.real)
        Zp[M - k] = complex(Ze.real + G.imag, -Ze.imag + G.real)
    Zp[M // 2] = 2.0 * np.conj(Xs[M // 2])
    z = np.fft.ifft(Zp) * M
    back = np.empty(nx)
    back[0::2], back[1::2] = z.real, z.imag
    assert np.abs(back - nx * x).max() < 1e-13 * nx
