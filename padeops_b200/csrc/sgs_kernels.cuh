// sgs_kernels.cuh — pointwise kernels of the eddy-viscosity SGS models (sgs_models/{smagorinsky, sigma, AMD, eddyViscosity}.F90)
// as __host__ __device__ functions: the CUDA kernels of igrid.cu and a host-only test hook run the same code.
// d[9]: dudx dudy dudz dvdx dvdy dvdz dwdx dwdy dwdz (the reference's duidxj(:,:,:,1:9)); S[6]: S11 S12 S13 S22 S23 S33.
#pragma once
#include <cmath>

#include "nonperiodic.cuh"   // PDO_HD

namespace pdo {

struct SgsConst {
    int mid;            // 0 Smagorinsky, 1 sigma, 2 AMD
    double cmodel;      // cmodel_global: (Cs deltaLES)^2 for 0 and 1, one for AMD
    double cx, cy, cz;  // AMD: camd_x, camd_y, camd_z
};

PDO_HD void sgs_sij(const double* d, double* S) {   // eddyViscosity.F90:1-24
    S[0] = d[0];
    S[3] = d[4];
    S[5] = d[8];
    S[1] = 0.5 * (d[1] + d[3]);
    S[2] = 0.5 * (d[2] + d[6]);
    S[4] = 0.5 * (d[5] + d[7]);
}

// the model's kernel (before the model constant)
PDO_HD double sgs_kernel_point(const SgsConst& c, const double* d, const double* S) {
    if (c.mid == 0) {   // smagorinsky.F90:44-66
        double t = S[0] * S[0];
        t = t + 2.0 * (S[1] * S[1]);
        t = t + 2.0 * (S[2] * S[2]);
        t = t + (S[3] * S[3]);
        t = t + 2.0 * (S[4] * S[4]);
        t = t + (S[5] * S[5]);
        t = 2.0 * t;
        return sqrt(t);
    }
    if (c.mid == 1) {   // sigma.F90:31-98
        const double kPiLocal = 3.141592653589793238462643383279502884197;
        const double G11 = d[0] * d[0] + d[3] * d[3] + d[6] * d[6];
        const double G12 = d[0] * d[1] + d[3] * d[4] + d[6] * d[7];
        const double G13 = d[0] * d[2] + d[3] * d[5] + d[6] * d[8];
        const double G22 = d[1] * d[1] + d[4] * d[4] + d[7] * d[7];
        const double G23 = d[1] * d[2] + d[4] * d[5] + d[7] * d[8];
        const double G33 = d[2] * d[2] + d[5] * d[5] + d[8] * d[8];
        const double I1 = G11 + G22 + G33;
        const double I1sq = I1 * I1;
        const double I1cu = I1sq * I1;
        double I2 = -G11 * G11 - G22 * G22 - G33 * G33;
        I2 = I2 - 2.0 * G12 * G12 - 2.0 * G13 * G13;
        I2 = I2 - 2.0 * G23 * G23;
        I2 = I2 + I1sq;
        I2 = 0.5 * I2;
        double I3 = G11 * (G22 * G33 - G23 * G23);
        I3 = I3 + G12 * (G13 * G23 - G12 * G33);
        I3 = I3 + G13 * (G12 * G23 - G22 * G13);
        double alpha1 = I1sq / 9.0 - I2 / 3.0;
        alpha1 = fmax(alpha1, 0.0);
        const double alpha2 = I1cu / 27.0 - I1 * I2 / 6.0 + I3 / 2.0;
        const double a1s = sqrt(alpha1);
        double t = alpha1 * a1s;
        t = alpha2 / (t + 1.0e-13);
        t = fmin(t, 1.0);
        t = fmax(t, -1.0);
        t = acos(t);
        const double alpha3 = (1.0 / 3.0) * t;
        double s1sq = I1 / 3.0 + 2.0 * a1s * cos(alpha3);
        s1sq = fmax(s1sq, 0.0);
        const double s1 = sqrt(s1sq);
        double s2 = kPiLocal / 3.0 + alpha3;
        s2 = (-2.0) * a1s * cos(s2);
        s2 = s2 + I1 / 3.0;
        s2 = sqrt(fmax(s2, 0.0));
        double s3 = kPiLocal / 3.0 - alpha3;
        s3 = (-2.0) * a1s * cos(s3);
        s3 = s3 + I1 / 3.0;
        s3 = sqrt(fmax(s3, 0.0));
        return s3 * (s1 - s2) * (s2 - s3) / (s1sq + 1.0e-15);
    }
    // AMD.F90:24-72
    const double cx = c.cx, cy = c.cy, cz = c.cz;
#define PDO_AMD_ROW(a, b) ((d[a] * cx) * (d[b] * cx) + (d[(a) + 1] * cy) * (d[(b) + 1] * cy) + (d[(a) + 2] * cz) * (d[(b) + 2] * cz))
    double num = PDO_AMD_ROW(0, 0) * S[0];
    num = num + PDO_AMD_ROW(3, 3) * S[3];
    num = num + PDO_AMD_ROW(6, 6) * S[5];
    num = num + 2.0 * PDO_AMD_ROW(0, 3) * S[1];
    num = num + 2.0 * PDO_AMD_ROW(0, 6) * S[2];
    num = num + 2.0 * PDO_AMD_ROW(3, 6) * S[4];
#undef PDO_AMD_ROW
    const double den = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5] + d[6] * d[6] + d[7] * d[7] + d[8] * d[8];
    return fmax(-num / (den + 1.0e-32), 0.0);
}

}  // namespace pdo
