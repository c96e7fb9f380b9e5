// handles.h — what the opaque operator handles of include/padeops_b200.h point to (shared by capi_ops.cu and distops.cu).
#pragma once
#include "banded.cuh"
#include "nonperiodic.cuh"
#include "stagg_np.cuh"
#include "../../include/padeops_b200.h"

struct pdo_cd10_s { int n; pdo::BandedOp d1, d2; bool periodic = true; pdo::NpOp np_d1, np_d2; };   // np*: periodic = .false. closures
struct pdo_cd06_s { int n; pdo::BandedOp d1; bool periodic = true; pdo::NpOp np; };
struct pdo_cf90_s { int n; pdo::BandedOp op; bool periodic = true; pdo::NpOp np; };
struct pdo_gaussian_s { int n; pdo::BandedOp op; bool periodic = true; pdo::NpOp np; };
struct pdo_lstsq_s { int n; pdo::BandedOp op; bool periodic = true; pdo::NpOp np; };
struct pdo_cd06stagg_s { int n; pdo::BandedOp ops[6]; bool periodic = true; pdo::StaggNp np; };   // np: init_nonperiodic (walls)
struct pdo_derivatives_s {
    int xsz[3], ysz[3], zsz[3];
    int method[3];  // 0 cd10, 1 cd06
    pdo_cd10_t c10[3];
    pdo_cd06_t c06[3];
};
struct pdo_filters_s {
    int xsz[3], ysz[3], zsz[3];
    int method[3];  // 0 cf90, 1 gaussian, 2 lstsq
    pdo_cf90_t cf[3];
    pdo_gaussian_t ga[3];
    pdo_lstsq_t ls[3];
};
