"""CPU restatement of the NON-PERIODIC staggered 6th-order compact operators (cd06stagg%init_nonperiodic and its eight
z-operators) — TEST INFRASTRUCTURE ONLY; groundwork for SURVEY.md §8f rank 2 (wall-bounded igrid), the CUDA library does not
implement these yet.

Follows derivatives/cd06stagg.F90:17-56 (constants), 197-231 (init_nonperiodic), 820-1059 (the operators) and the included
derivatives/STAGG_CD06_files/*.F90 statement by statement:
    ComputeTri_allRoutines.F90   the eight tridiagonal systems and their Thomas factors (ddn*den, den, cp)
    TridiagSolver_allRoutines.F90:1-17   SolveZTriREAL / CMPLX
    D1RHS_{E2C,C2E,C2C,E2E}_common.F90, InterpRHS_{E2C,C2E}_common.F90, D2RHS_{C2C,E2E}_common.F90
Fields are numpy arrays in the repo's view of the Fortran layout: f(n1,n2,n) has shape (n, n2, n1), so the staggered (z)
index is axis 0 and every statement below is vectorised over the (n2, n1) planes exactly like the Fortran array syntax.
Cells sit at (k-1/2) dz, edges at (k-1) dz, k = 1.., nE = n + 1.  isBotEven / isTopEven: the FIELD is even (True) or odd
(False) about that wall; isBotSided / isTopSided: one-sided closure at that wall instead (no symmetry assumed).
Real and complex fields take the same code (the Fortran includes the same file into the REAL and CMPLX procedures)."""
import numpy as np

# ---- constants (cd06stagg.F90:17-56) ----
alpha, p, q, r, s = 3.0, 17.0 / 6.0, 3.0 / 2.0, 3.0 / 2.0, -1.0 / 6.0
alpha06d1, a06d1, b06d1 = 1.0 / 3.0, (14.0 / 9.0) / 2.0, (1.0 / 9.0) / 4.0
qhat, rhat, alpha_hat = a06d1, b06d1, alpha06d1
q_p, alpha_p = 3.0 / 4.0, 1.0 / 4.0
alpha_pp = ((40 * alpha_hat - 1) * q + 7 * (4 * alpha_hat - 1) * s) / (16 * (alpha_hat + 2) * q + 8 * (1 - 4 * alpha_hat) * s)
q_pp = (1.0 / 3.0) * (alpha_pp + 2)
r_pp = (1.0 / 12.0) * (4 * alpha_pp - 1)
w1 = (2 * alpha_hat + 1) / (2 * (q + s))
w2 = ((8 * alpha_hat + 7) * q - 6 * (2 * alpha_hat + 1) * r + (8 * alpha_hat + 7) * s) / (9 * (q + s))
w3 = (4 * (alpha_hat + 2) * q + 2 * (1 - 4 * alpha_hat) * s) / (9 * (q + s))
w0s, w1s = 223.0 / 186.0, 61.0 / 62.0


def _factor(ddn, dg, dup):
    """cp / den recurrences shared by every ComputeTri* routine; returns Tri(n,3) columns (ddn*den, den, cp) and the raw rows."""
    n = dg.size
    cp = np.zeros(n)
    den = np.zeros(n)
    cp[0] = dup[0] / dg[0]
    for i in range(1, n - 1):
        cp[i] = dup[i] / (dg[i] - ddn[i] * cp[i - 1])
    den[0] = 1.0 / dg[0]
    den[1:] = 1.0 / (dg[1:] - ddn[1:] * cp[:-1])
    return {"t1": ddn * den, "t2": den, "t3": cp, "rows": (ddn.copy(), dg.copy(), dup.copy())}


def _solve(T, y):
    """SolveZTriREAL / SolveZTriCMPLX (TridiagSolver_allRoutines.F90), in place on a copy; z is axis 0."""
    y = y.copy()
    n = y.shape[0]
    t1, t2, t3 = T["t1"], T["t2"], T["t3"]
    y[0] = y[0] * t2[0]
    for k in range(1, n):
        y[k] = y[k] * t2[k] - y[k - 1] * t1[k]
    for k in range(n - 2, -1, -1):
        y[k] = y[k] - t3[k] * y[k + 1]
    return y


class CD06StaggNP:
    def __init__(self, nx, dx, isTopEven, isBotEven, isTopSided=False, isBotSided=False):
        if nx <= 4:
            raise ValueError("CD06_stagg requires at least 4 points (code 21)")   # cd06stagg.F90:216-218
        self.n, self.nE = nx, nx + 1
        self.dx, self.onebydx = dx, 1.0 / dx
        self.onebydx2 = self.onebydx / dx
        self.isTopEven, self.isBotEven, self.isTopSided, self.isBotSided = isTopEven, isBotEven, isTopSided, isBotSided
        self.TriD1_E2C = self._tri_d1_e2c()
        self.TriD1_C2E = self._tri_d1_c2e()
        self.TriD1_E2E = self._tri_d1_e2e()
        self.TriD1_C2C = self._tri_d1_c2c()
        self.TriD2_E2E = self._tri_d2_e2e()
        self.TriD2_C2C = self._tri_d2_c2c()
        self.TriInterp_E2C = self._tri_interp_e2c()
        self.TriInterp_C2E = self._tri_interp_c2e()

    # ------------------------------------------------------------- ComputeTri_allRoutines.F90
    def _tri_d1_e2c(self):          # :1-47
        al, al1, al0 = 9.0 / 62.0, 37.0 / 183.0, -1.0
        n = self.n
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopSided:
            ddn[n - 1] = w0s * al0; dg[n - 1] = w0s
            ddn[n - 2] = w1s * al1; dup[n - 2] = w1s * al1; dg[n - 2] = w1s
        elif self.isTopEven:
            dup[n - 1] = 0.0; dg[n - 1] = 1.0 - al
        else:
            dg[n - 1] = 1.0 + al; dup[n - 1] = 0.0
        if self.isBotSided:
            dup[0] = w0s * al0; dg[0] = w0s
            ddn[1] = w1s * al1; dup[1] = w1s * al1; dg[1] = w1s
        elif self.isBotEven:
            ddn[0] = 0.0; dg[0] = 1.0 - al
        else:
            ddn[0] = 0.0; dg[0] = 1.0 + al
        return _factor(ddn, dg, dup)

    def _tri_d1_c2e(self):          # :49-92
        al = 9.0 / 62.0
        n = self.nE
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopSided:
            ddn[n - 1] = 0.0; dup[n - 1] = 0.0
            ddn[n - 2] = 1.0 / 22.0; dup[n - 2] = 1.0 / 22.0
        elif self.isTopEven:
            dup[n - 1] = 0.0; ddn[n - 1] = 0.0
        else:
            ddn[n - 1] = 2.0 * al; dup[n - 1] = 0.0
        if self.isBotSided:
            dup[0] = 0.0; ddn[0] = 0.0
            dup[1] = 1.0 / 22.0; ddn[1] = 1.0 / 22.0
        elif self.isBotEven:
            ddn[0] = 0.0; dup[0] = 0.0
        else:
            ddn[0] = 0.0; dup[0] = 2.0 * al
        return _factor(ddn, dg, dup)

    def _tri_d1_c2c(self):          # :94-160
        al, alLOW = 1.0 / 3.0, 3.0
        n = self.n
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopSided:
            dup[n - 1] = w1 * 0.0; dup[n - 2] = w2 * alpha_p; dup[n - 3] = w3 * alpha_pp
            dg[n - 1] = w1 * 1.0; dg[n - 2] = w2 * 1.0; dg[n - 3] = w3 * 1.0
            ddn[n - 1] = w1 * alLOW; ddn[n - 2] = w2 * alpha_p; ddn[n - 3] = w3 * alpha_pp
        elif self.isTopEven:
            dg[n - 1] = 1.0 - al
        else:
            dg[n - 1] = 1.0 + al
        if self.isBotSided:
            ddn[0] = w1 * 0.0; ddn[1] = w2 * alpha_p; ddn[2] = w3 * alpha_pp
            dg[0] = w1 * 1.0; dg[1] = w2 * 1.0; dg[2] = w3 * 1.0
            dup[0] = w1 * alLOW; dup[1] = w2 * alpha_p; dup[2] = w3 * alpha_pp
        elif self.isBotEven:
            dg[0] = 1.0 - al
        else:
            dg[0] = 1.0 + al
        return _factor(ddn, dg, dup)

    def _tri_d1_e2e(self):          # :162-195
        al = 1.0 / 3.0
        n = self.nE
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopEven:
            dg[n - 1] = 1.0; ddn[n - 1] = 0.0
        else:
            dg[n - 1] = 1.0; ddn[n - 1] = 2.0 * al
        if self.isBotEven:
            dg[0] = 1.0; dup[0] = 0.0
        else:
            dg[0] = 1.0; dup[0] = 2.0 * al
        return _factor(ddn, dg, dup)

    def _tri_interp_c2e(self):      # :197-241
        al, al1 = 3.0 / 10.0, 1.0 / 6.0
        n = self.nE
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopSided:
            dup[n - 1] = 0.0; ddn[n - 1] = 0.0
            dup[n - 2] = al1; ddn[n - 2] = al1
        elif self.isTopEven:
            ddn[n - 1] = 2.0 * al; dg[n - 1] = 1.0
        else:
            ddn[n - 1] = 0.0; dg[n - 1] = 1.0
        if self.isBotSided:
            dup[0] = 0.0; ddn[0] = 0.0; dg[0] = 1.0
            dup[1] = al1; ddn[1] = al1; dg[1] = 1.0
        elif self.isBotEven:
            dup[0] = 2.0 * al; dg[0] = 1.0
        else:
            dup[0] = 0.0; dg[0] = 1.0
        return _factor(ddn, dg, dup)

    def _tri_interp_e2c(self):      # :244-287
        al, al0 = 3.0 / 10.0, 1.0
        n = self.n
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopSided:
            ddn[n - 1] = al0; dup[n - 1] = 0.0
        elif self.isTopEven:
            dg[n - 1] = 1.0 + al
        else:
            dg[n - 1] = 1.0 - al
        if self.isBotSided:
            dup[0] = al0; ddn[0] = 0.0
        elif self.isBotEven:
            dg[0] = 1.0 + al
        else:
            dg[0] = 1.0 - al
        return _factor(ddn, dg, dup)

    def _tri_d2_e2e(self):          # :290-325
        al = 2.0 / 11.0
        n = self.nE
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        if self.isTopEven:
            dg[n - 1] = 1.0; ddn[n - 1] = 2.0 * al
        else:
            dg[n - 1] = 1.0; ddn[n - 1] = 0.0
        if self.isBotEven:
            dg[0] = 1.0; dup[0] = 2.0 * al
        else:
            dg[0] = 1.0; dup[0] = 0.0
        return _factor(ddn, dg, dup)

    def _tri_d2_c2c(self):          # :327-366
        al = 2.0 / 11.0
        n = self.n
        ddn, dg, dup = np.full(n, al), np.ones(n), np.full(n, al)
        dg[n - 1] = 1.0 + al if self.isTopEven else 1.0 - al
        dg[0] = 1.0 + al if self.isBotEven else 1.0 - al
        return _factor(ddn, dg, dup)

    # ------------------------------------------------------------- right-hand sides (0-based: Fortran index k -> k-1)
    def _rhs_d1_e2c(self, fE):      # D1RHS_E2C_common.F90
        n, nE, o = self.n, self.nE, self.onebydx
        a, b = (63.0 / 62.0) / 1.0, (17.0 / 62.0) / 3.0
        al1, al0 = 37.0 / 183.0, -1.0
        a0, b0 = (1.0 / 24.0) * (al0 - 23.0), (1.0 / 8.0) * (-9.0 * al0 + 7.0)
        c0, d0 = (1.0 / 8.0) * (9.0 * al0 + 1.0), -(1.0 / 24.0) * (al0 + 1.0)
        a1, b1 = (3.0 / 8.0) * (3.0 - 2.0 * al1), (1.0 / 8.0) * (-1.0 + 22.0 * al1)
        a06, b06 = a * o, b * o
        rhs = np.zeros((n,) + fE.shape[1:], dtype=fE.dtype)
        rhs[1:n - 1] = b06 * (fE[3:nE] - fE[0:nE - 3]) + a06 * (fE[2:nE - 1] - fE[1:nE - 2])
        if self.isBotSided:
            rhs[0] = w0s * o * (a0 * fE[0] + b0 * fE[1] + c0 * fE[2] + d0 * fE[3])
            rhs[1] = w1s * o * ((-b1 / 3.0) * fE[0] + (-a1) * fE[1] + (a1) * fE[2] + (b1 / 3.0) * fE[3])
        elif self.isBotEven:
            rhs[0] = b06 * (fE[2] - fE[1]) + a06 * (fE[1] - fE[0])
        else:
            rhs[0] = b06 * (fE[2] + fE[1]) + a06 * (fE[1] - fE[0])
        if self.isTopSided:
            rhs[n - 1] = -w0s * o * (a0 * fE[nE - 1] + b0 * fE[nE - 2] + c0 * fE[nE - 3] + d0 * fE[nE - 4])
            rhs[n - 2] = w1s * o * ((-b1 / 3.0) * fE[nE - 4] + (-a1) * fE[nE - 3] + (a1) * fE[nE - 2] + (b1 / 3.0) * fE[nE - 1])
        elif self.isTopEven:
            rhs[n - 1] = b06 * (fE[nE - 2] - fE[nE - 3]) + a06 * (fE[nE - 1] - fE[nE - 2])
        else:
            rhs[n - 1] = -b06 * (fE[nE - 2] + fE[nE - 3]) + a06 * (fE[nE - 1] - fE[nE - 2])
        return rhs

    def _rhs_d1_c2e(self, fC):      # D1RHS_C2E_common.F90
        n, nE, o = self.n, self.nE, self.onebydx
        a, b = (63.0 / 62.0) / 1.0, (17.0 / 62.0) / 3.0
        a0, b0, c0, d0 = -71.0 / 24.0, 47.0 / 8.0, -31.0 / 8.0, 23.0 / 24.0
        a1 = 12.0 / 11.0
        a06, b06 = a * o, b * o
        rhs = np.zeros((nE,) + fC.shape[1:], dtype=fC.dtype)
        rhs[1:nE - 1] = a06 * (fC[1:n] - fC[0:n - 1])
        rhs[2:nE - 2] = rhs[2:nE - 2] + b06 * (fC[3:n] - fC[0:n - 3])
        if self.isBotSided:
            rhs[0] = (a0 * fC[0] + b0 * fC[1] + c0 * fC[2] + d0 * fC[3]) * o
            rhs[1] = (fC[1] - fC[0]) * (a1 * o)
        elif self.isBotEven:
            rhs[0] = 0.0
            rhs[1] = rhs[1] + b06 * (fC[2] - fC[0])
        else:
            rhs[0] = 2.0 * b06 * fC[1] + 2.0 * a06 * fC[0]
            rhs[1] = rhs[1] + b06 * (fC[2] + fC[0])
        if self.isTopSided:
            rhs[nE - 2] = (fC[n - 1] - fC[n - 2]) * (a1 * o)
            rhs[nE - 1] = (a0 * fC[n - 1] + b0 * fC[n - 2] + c0 * fC[n - 3] + d0 * fC[n - 4]) * (-o)
        elif self.isTopEven:
            rhs[nE - 2] = rhs[nE - 2] + b06 * (fC[n - 1] - fC[n - 3])
            rhs[nE - 1] = 0.0
        else:
            rhs[nE - 2] = rhs[nE - 2] - b06 * (fC[n - 1] + fC[n - 3])
            rhs[nE - 1] = -2.0 * b06 * (fC[n - 2]) - 2.0 * a06 * (fC[n - 1])
        return rhs

    def _rhs_d1_c2c(self, fC):      # D1RHS_C2C_common.F90
        n, o = self.n, self.onebydx
        a, b = (14.0 / 9.0) / 2.0, (1.0 / 9.0) / 4.0
        a06, b06 = a * o, b * o
        a_np_3, b_np_3 = w3 * q_pp * o, w3 * r_pp * o
        a_np_2 = w2 * q_p * o
        a_np_1, b_np_1, c_np_1, d_np_1 = w1 * (-p * o), w1 * (q * o), w1 * (r * o), w1 * (s * o)
        rhs = np.zeros_like(fC)
        if self.isBotSided:
            rhs[0] = a_np_1 * fC[0] + b_np_1 * fC[1] + c_np_1 * fC[2] + d_np_1 * fC[3]
            rhs[1] = a_np_2 * (fC[2] - fC[0])
            rhs[2] = a_np_3 * (fC[3] - fC[1]) + b_np_3 * (fC[4] - fC[0])
        elif self.isBotEven:
            rhs[0] = b06 * (fC[2] - fC[1]) + a06 * (fC[1] - fC[0])
            rhs[1] = b06 * (fC[3] - fC[0]) + a06 * (fC[2] - fC[0])
            rhs[2] = b06 * (fC[4] - fC[0]) + a06 * (fC[3] - fC[1])
        else:
            rhs[0] = b06 * (fC[2] + fC[1]) + a06 * (fC[1] + fC[0])
            rhs[1] = b06 * (fC[3] + fC[0]) + a06 * (fC[2] - fC[0])
            rhs[2] = b06 * (fC[4] - fC[0]) + a06 * (fC[3] - fC[1])
        rhs[3:n - 3] = b06 * (fC[5:n - 1] - fC[1:n - 5]) + a06 * (fC[4:n - 2] - fC[2:n - 4])
        if self.isTopSided:
            rhs[n - 3] = a_np_3 * (fC[n - 2] - fC[n - 4]) + b_np_3 * (fC[n - 1] - fC[n - 5])
            rhs[n - 2] = a_np_2 * (fC[n - 1] - fC[n - 3])
            rhs[n - 1] = -a_np_1 * fC[n - 1] - b_np_1 * fC[n - 2] - c_np_1 * fC[n - 3] - d_np_1 * fC[n - 4]
        elif self.isTopEven:
            rhs[n - 3] = b06 * (fC[n - 1] - fC[n - 5]) + a06 * (fC[n - 2] - fC[n - 4])
            rhs[n - 2] = b06 * (fC[n - 1] - fC[n - 4]) + a06 * (fC[n - 1] - fC[n - 3])
            rhs[n - 1] = b06 * (fC[n - 2] - fC[n - 3]) + a06 * (fC[n - 1] - fC[n - 2])
        else:
            rhs[n - 3] = b06 * (fC[n - 1] - fC[n - 5]) + a06 * (fC[n - 2] - fC[n - 4])
            rhs[n - 2] = -b06 * (fC[n - 1] + fC[n - 4]) + a06 * (fC[n - 1] - fC[n - 3])
            rhs[n - 1] = -b06 * (fC[n - 2] + fC[n - 3]) - a06 * (fC[n - 1] + fC[n - 2])
        return rhs

    def _rhs_d1_e2e(self, fE):      # D1RHS_E2E_common.F90
        nE, o = self.nE, self.onebydx
        a, b = (14.0 / 9.0) / 2.0, (1.0 / 9.0) / 4.0
        a06, b06 = a * o, b * o
        rhs = np.zeros_like(fE)
        rhs[1:nE - 1] = a06 * (fE[2:nE] - fE[0:nE - 2])
        rhs[2:nE - 2] = rhs[2:nE - 2] + b06 * (fE[4:nE] - fE[0:nE - 4])
        if self.isBotEven:
            rhs[0] = 0.0
            rhs[1] = rhs[1] + b06 * (fE[3] - fE[1])
        else:
            rhs[0] = b06 * (fE[2] + fE[2]) + a06 * (fE[1] + fE[1])
            rhs[1] = rhs[1] + b06 * (fE[3] + fE[1])
        if self.isTopEven:
            rhs[nE - 1] = 0.0
            rhs[nE - 2] = rhs[nE - 2] + b06 * (fE[nE - 2] - fE[nE - 4])
        else:
            rhs[nE - 1] = -b06 * (fE[nE - 3] + fE[nE - 3]) - a06 * (fE[nE - 2] + fE[nE - 2])
            rhs[nE - 2] = rhs[nE - 2] - b06 * (fE[nE - 2] + fE[nE - 4])
        return rhs

    def _rhs_interp_e2c(self, fE):  # InterpRHS_E2C_common.F90
        n, nE = self.n, self.nE
        b, a = (1.0 / 10.0) / 2.0, (3.0 / 2.0) / 2.0
        al0 = 1.0
        a0, b0 = (1.0 / 16.0) * (5.0 - al0), (1.0 / 16.0) * (9.0 * al0 + 15.0)
        c0, d0 = (1.0 / 16.0) * (9.0 * al0 - 5.0), (1.0 / 16.0) * (1.0 - al0)
        rhs = np.zeros((n,) + fE.shape[1:], dtype=fE.dtype)
        if self.isBotSided:
            rhs[0] = a0 * fE[0] + b0 * fE[1] + c0 * fE[2] + d0 * fE[3]
        elif self.isBotEven:
            rhs[0] = (b) * (fE[2] + fE[1]) + (a) * (fE[1] + fE[0])
        else:
            rhs[0] = (b) * (fE[2] - fE[1]) + (a) * (fE[1] + fE[0])
        rhs[1:n - 1] = (b) * (fE[3:nE] + fE[0:nE - 3]) + (a) * (fE[2:nE - 1] + fE[1:nE - 2])
        if self.isTopSided:
            rhs[n - 1] = a0 * fE[nE - 1] + b0 * fE[nE - 2] + c0 * fE[nE - 3] + d0 * fE[nE - 4]
        elif self.isTopEven:
            rhs[n - 1] = (b) * (fE[nE - 2] + fE[nE - 3]) + (a) * (fE[nE - 1] + fE[nE - 2])
        else:
            rhs[n - 1] = (b) * (-fE[nE - 2] + fE[nE - 3]) + (a) * (fE[nE - 1] + fE[nE - 2])
        return rhs

    def _rhs_interp_c2e(self, fC):  # InterpRHS_C2E_common.F90
        n, nE = self.n, self.nE
        b, a = (1.0 / 10.0) / 2.0, (3.0 / 2.0) / 2.0
        a0, b0, c0 = 15.0 / 8.0, -5.0 / 4.0, 3.0 / 8.0
        al1 = 1.0 / 6.0
        a1 = (1.0 / 8.0) * (9.0 + 10.0 * al1)
        rhs = np.zeros((nE,) + fC.shape[1:], dtype=fC.dtype)
        rhs[1:nE - 1] = (a) * (fC[1:n] + fC[0:n - 1])
        rhs[2:nE - 2] = rhs[2:nE - 2] + (b) * (fC[3:n] + fC[0:n - 3])
        if self.isBotSided:
            rhs[0] = a0 * fC[0] + b0 * fC[1] + c0 * fC[2]
            rhs[1] = (a1 / 2.0) * (fC[1] + fC[0])
        elif self.isBotEven:
            rhs[0] = 2.0 * b * fC[1] + 2.0 * a * fC[0]
            rhs[1] = rhs[1] + (b) * (fC[2] + fC[0])
        else:
            rhs[0] = 0.0
            rhs[1] = rhs[1] + (b) * (fC[2] - fC[0])
        if self.isTopSided:
            rhs[nE - 1] = a0 * fC[n - 1] + b0 * fC[n - 2] + c0 * fC[n - 3]
            rhs[nE - 2] = (a1 / 2.0) * (fC[n - 1] + fC[n - 2])
        elif self.isTopEven:
            rhs[nE - 1] = 2.0 * b * fC[n - 2] + 2.0 * a * fC[n - 1]
            rhs[nE - 2] = rhs[nE - 2] + (b) * (fC[n - 1] + fC[n - 3])
        else:
            rhs[nE - 1] = 0.0
            rhs[nE - 2] = rhs[nE - 2] + (b) * (-fC[n - 1] + fC[n - 3])
        return rhs

    def _rhs_d2_c2c(self, fC):      # D2RHS_C2C_common.F90
        n = self.n
        a, b = 12.0 / 11.0, (3.0 / 11.0) / 4.0
        a06, b06 = a * self.onebydx2, b * self.onebydx2
        rhs = np.zeros_like(fC)
        rhs[1:n - 1] = a06 * (fC[2:n] + fC[0:n - 2])
        rhs[2:n - 2] = rhs[2:n - 2] + b06 * (fC[4:n] + fC[0:n - 4])
        if self.isBotEven:
            rhs[0] = b06 * (fC[2] + fC[1]) + a06 * (fC[1] + fC[0])
            rhs[1] = rhs[1] + b06 * (fC[3] + fC[0])
        else:
            rhs[0] = b06 * (fC[2] - fC[1]) + a06 * (fC[1] - fC[0])
            rhs[1] = rhs[1] + b06 * (fC[3] - fC[0])
        if self.isTopEven:
            rhs[n - 1] = b06 * (fC[n - 2] + fC[n - 3]) + a06 * (fC[n - 1] + fC[n - 2])
            rhs[n - 2] = rhs[n - 2] + b06 * (fC[n - 1] + fC[n - 4])
        else:
            rhs[n - 1] = b06 * (-fC[n - 2] + fC[n - 3]) + a06 * (-fC[n - 1] + fC[n - 2])
            rhs[n - 2] = rhs[n - 2] + b06 * (-fC[n - 1] + fC[n - 4])
        return rhs - 2.0 * (b06 + a06) * fC

    def _rhs_d2_e2e(self, fE):      # D2RHS_E2E_common.F90
        nE = self.nE
        a, b = 12.0 / 11.0, (3.0 / 11.0) / 4.0
        a06, b06 = a * self.onebydx2, b * self.onebydx2
        rhs = np.zeros_like(fE)
        rhs[1:nE - 1] = a06 * (fE[2:nE] + fE[0:nE - 2])
        rhs[2:nE - 2] = rhs[2:nE - 2] + b06 * (fE[4:nE] + fE[0:nE - 4])
        if self.isBotEven:
            rhs[0] = b06 * (fE[2] + fE[2]) + a06 * (fE[1] + fE[1])
            rhs[1] = rhs[1] + b06 * (fE[3] + fE[1])
        else:
            rhs[0] = 0.0
            rhs[1] = rhs[1] + b06 * (fE[3] - fE[1])
        if self.isTopEven:
            rhs[nE - 1] = 2.0 * b06 * fE[nE - 3] + 2.0 * a06 * fE[nE - 2]
            rhs[nE - 2] = rhs[nE - 2] + b06 * (fE[nE - 2] + fE[nE - 4])
        else:
            rhs[nE - 1] = 0.0
            rhs[nE - 2] = rhs[nE - 2] + b06 * (-fE[nE - 2] + fE[nE - 4])
        return rhs - 2.0 * (b06 + a06) * fE

    # ------------------------------------------------------------- the operators (cd06stagg.F90:820-1059, non-periodic branch)
    def ddz_E2C(self, fE):
        return _solve(self.TriD1_E2C, self._rhs_d1_e2c(np.asarray(fE)))

    def ddz_C2E(self, fC):
        return _solve(self.TriD1_C2E, self._rhs_d1_c2e(np.asarray(fC)))

    def ddz_C2C(self, fC):
        return _solve(self.TriD1_C2C, self._rhs_d1_c2c(np.asarray(fC)))

    def ddz_E2E(self, fE):
        return _solve(self.TriD1_E2E, self._rhs_d1_e2e(np.asarray(fE)))

    def InterpZ_E2C(self, fE):
        return _solve(self.TriInterp_E2C, self._rhs_interp_e2c(np.asarray(fE)))

    def InterpZ_C2E(self, fC):
        return _solve(self.TriInterp_C2E, self._rhs_interp_c2e(np.asarray(fC)))

    def d2dz2_C2C(self, fC):
        return _solve(self.TriD2_C2C, self._rhs_d2_c2c(np.asarray(fC)))

    def d2dz2_E2E(self, fE):
        return _solve(self.TriD2_E2E, self._rhs_d2_e2e(np.asarray(fE)))
