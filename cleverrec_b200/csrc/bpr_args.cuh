// Arguments of the fused BPR step kernels (train.cu: register-staged gathers; train_ring.cu: bulk-copy ring).
#pragma once
#include "rowopt.cuh"

struct BprArgs {
    TableDev P, Q;
    unsigned long long* metaU;
    unsigned long long* metaI;
    const int32_t* u;
    const int32_t* i;
    const int32_t* j;
    const uint32_t* rk[3];
    int64_t batch;
    int dim;
    float reg;
    OptDev opt;
    float* dup_grad;
    uint32_t* dup_t;
    double* block_loss;
};

// train_ring.cu: the same step with the row gathers issued as cp.async.bulk copies into a shared-memory ring (mbarrier hand-off)
int crb_launch_bpr_ring(crb_handle* h, const BprArgs& a, int opt_kind, cudaStream_t s);
bool crb_bpr_ring_enabled(int dim);
