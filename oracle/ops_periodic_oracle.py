"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of igrid_Operators_Periodic::Ops_Periodic
(src/incompressible/igrid_operators_periodic.F90:13-161), the triply periodic post-processing operators — SURVEY.md 8f rank 1.
Never imported by the product.

  init           :86-109   spectral("x", nx, ny, nz, "four", "2/3rd", dimTransform 2, fixOddball = .false., init_periodicInZ,
                           dealiasF = 2/3) + PoissonPeriodic(dir_id = 1) with Fourier wavenumbers on all three axes
  ddx / ddy      :117-135  fft (2-D, x and y), x i k1 / i k2, ifft
  ddz            :149-160  x->y->z transposes, spectral%ddz_C2C_real_inplace (spectral.F90:507-526), back
  ddz_cmplx2cmplx:137-145  y->z, spectral%ddz_C2C_complex_inplace (:528-547), z->y
  SolvePoisson   :70-84    PoissonPeriodic%poisson_solve
  dealiasField   :56-62    fft, spectral%dealias (3-D 2/3 box, spectral.F90:785-814), ifft

Arrays are global, C order (nz, ny, nx) [complex: (nz, ny, nx/2+1)]; the transposes are pure permutations."""
import numpy as np

from . import igrid_oracle as IG
from . import oracle as O


class OpsPeriodic:
    def __init__(self, nx, ny, nz, dx, dy, dz):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dx, self.dy, self.dz = dx, dy, dz
        self.spect = IG.Spectral(nx, ny, nz, dx, dy, dz, init_periodicInZ=True, dealiasF=2.0 / 3.0, fixOddball=False)

    def ddx(self, f):
        return self.spect.ifft(self.spect.mTimes_ik1(self.spect.fft(f)))

    def ddy(self, f):
        return self.spect.ifft(self.spect.mTimes_ik2(self.spect.fft(f)))

    def ddz(self, f):
        return self.spect.ddz_C2C_real_inplace(f)

    def ddz_cmplx2cmplx(self, fhat):
        return self.spect.ddz_C2C_complex_inplace(fhat)

    def SolvePoisson(self, rhs):
        return O.poisson_solve(rhs, self.dx, self.dy, self.dz)

    def dealiasField(self, f):
        return self.spect.ifft(self.spect.dealias(self.spect.fft(f)))
