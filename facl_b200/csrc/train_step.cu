// One whole training step behind a single C-ABI call: the loop body of reference
// training_code/cn3d_train_motion_GL.py:224-335 (G-major flatten -> grouping -> encoder -> global + circle loss
// -> backward -> Adam), optionally fed from a pinned HOST batch and returning the loss to a pinned host float.
#include "../../include/facl_b200.h"
#include "common.cuh"
#include "facl_internal.h"

namespace facl {

namespace {
// (B,G,N,4) -> (G*B,N,4): cloud g*B+b = view g of sample b (cn3d_train_motion_GL.py:225-226); float4 rows
__global__ void gmajor_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int G, int N) {
    pdl_prologue();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * G * N;
    if (t >= total) return;
    int n = (int)(t % N);
    int m = (int)(t / N);
    int g = m / B, b = m % B;
    out[t] = in[((long long)b * G + g) * N + n];
}
__global__ void centres_kernel(const float* __restrict__ clouds, int M, int N, int S, float* __restrict__ centres) {
    pdl_prologue();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * S) return;
    int m = t / S, s = t % S;
    const float* p = clouds + ((long long)m * N + s) * 4;
    centres[t * 3 + 0] = p[0];
    centres[t * 3 + 1] = p[1];
    centres[t * 3 + 2] = p[2];
}
struct OrderVals { int v[FACL_MAX_VIEWS]; };
__global__ void set_order_kernel(const OrderVals o, int G, int* __restrict__ order) {
    pdl_prologue();
    if ((int)threadIdx.x < G) order[threadIdx.x] = o.v[threadIdx.x];
}
__global__ void add2_kernel(const float* __restrict__ v, float* __restrict__ out) {
    pdl_prologue(); out[0] = v[0] + v[1]; }
__global__ void axpy1_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
    pdl_prologue();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
}  // namespace

int gmajor_launch(const float* in, float* out, int B, int G, int N, cudaStream_t st) {
    ScopedTimer timer(TAG_TRANSPOSE, st);
    count_launch();
    long long total = (long long)B * G * N;
    FACL_LAUNCH_OK(launch_pdl(gmajor_kernel, dim3(div_up(total, 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), B, G, N));
    return (int)cudaGetLastError();
}
int centres_launch(const float* clouds, int M, int N, int S, float* centres, cudaStream_t st) {
    ScopedTimer timer(TAG_POOLMISC, st);
    count_launch();
    FACL_LAUNCH_OK(launch_pdl(centres_kernel, dim3(div_up((long long)M * S, 256)), dim3(256), 0, st, clouds, M, N, S, centres));
    return (int)cudaGetLastError();
}

}  // namespace facl

using namespace facl;

extern "C" {

int facl_gmajor(const float* points_bgnd, float* clouds, int B, int G, int N, void* stream) {
    return gmajor_launch(points_bgnd, clouds, B, G, N, reinterpret_cast<cudaStream_t>(stream));
}

int facl_train_step(const facl_train_step_args* a, void* stream) {
    if (!a || !a->dims || !a->params || !a->grads || !a->enc_buffers) return (int)cudaErrorInvalidValue;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const facl_encoder_dims* d = a->dims;
    const int M = d->M, G = d->G, Bl = M / G, S = d->S, K = d->K, N = a->N;
    const int phases = a->phases ? a->phases : FACL_PHASE_ALL;
    const int Bglob = a->B_global > 0 ? a->B_global : Bl;
    int rc;
    if (phases & (FACL_PHASE_FORWARD | FACL_PHASE_FORWARD_X)) {
        const float* batch = a->points_bgnd;
        if (a->points_host) {   // end-to-end path: the batch starts in pinned host memory
            FACL_CHECK(cudaMemcpyAsync(a->staging, a->points_host, sizeof(float) * 4 * (size_t)M * N, cudaMemcpyHostToDevice, st));
            batch = a->staging;
        }
        if (!batch) return (int)cudaErrorInvalidValue;
        if ((rc = gmajor_launch(batch, a->clouds, Bl, G, N, st))) return rc;
        if ((rc = group_launch(a->clouds, M, N, 4, S, K, a->r2, a->xt, nullptr, st))) return rc;
        if ((rc = centres_launch(a->clouds, M, N, S, a->centres, st))) return rc;
        if ((rc = encoder_forward(d, a->params, a->xt, a->centres, a->enc_buffers, a->x, a->x_global, nullptr, nullptr,
                                  (phases & FACL_PHASE_FORWARD) ? 3 : 1, st)))
            return rc;
    }
    if ((phases & FACL_PHASE_FORWARD_G) && !(phases & FACL_PHASE_FORWARD)) {
        if ((rc = encoder_forward(d, a->params, a->xt, a->centres, a->enc_buffers, a->x, a->x_global, nullptr, nullptr, 2, st))) return rc;
    }
    if (phases & FACL_PHASE_LOSS) {
        // single GPU: keys == x and both gradient roles are summed into dx; sharded: dkeys is a separate buffer that the
        // caller sum-reduce-scatters and hands back as dx_extra
        float* dkeys = a->keys ? a->dkeys : a->dx;
        if (a->order_by_value) {
            if (!a->order || G > FACL_MAX_VIEWS) return (int)cudaErrorInvalidValue;
            OrderVals ov;
            bool seen[FACL_MAX_VIEWS] = {false};
            for (int i = 0; i < G; ++i) {          // must be a permutation of 0..G-1: the loss kernels index rows with it
                const int o = a->order_vals[i];
                if (o < 0 || o >= G || seen[o]) return (int)cudaErrorInvalidValue;
                seen[o] = true;
                ov.v[i] = o;
            }
            count_launch();
            FACL_LAUNCH_OK(launch_pdl(set_order_kernel, dim3(1), dim3(FACL_MAX_VIEWS), 0, st, ov, G, a->order));
            FACL_CHECK_LAUNCH();
        }
        if ((rc = facl_contrast_losses(a->x, a->x_global, a->keys, G, Bglob, Bl, a->sample_offset, 512, a->order, 1, 1, 3 /* the losses run the split products in both modes: see facl_contrast_losses */,
                                       a->loss_ws, a->loss2, a->dx, a->dx_global, dkeys, stream)))
            return rc;
        count_launch();
        FACL_LAUNCH_OK(launch_pdl(add2_kernel, dim3(1), dim3(1), 0, st, a->loss2, a->loss2 + 2));
        FACL_CHECK_LAUNCH();
    }
    if ((phases & FACL_PHASE_BACKWARD_HEAD_G) && !(phases & (FACL_PHASE_BACKWARD | FACL_PHASE_BACKWARD_HEAD))) {
        // the sequence half of the head needs dx_global only: it runs before dx += dx_extra, beside the caller's reduce-scatter
        if ((rc = encoder_backward(d, a->params, a->xt, a->enc_buffers, nullptr, a->dx_global, a->grads, 4, st))) return rc;
    }
    if (phases & (FACL_PHASE_BACKWARD | FACL_PHASE_BACKWARD_HEAD | FACL_PHASE_BACKWARD_HEAD_X)) {
        if (a->dx_extra) {
            long long n = (long long)M * 512;
            count_launch();
            FACL_LAUNCH_OK(launch_pdl(axpy1_kernel, dim3(div_up(n, 256)), dim3(256), 0, st, a->dx, a->dx_extra, n));
            FACL_CHECK_LAUNCH();
        }
    }
    {
        // FACL_PHASE_BACKWARD = both stages; a sharded caller issues HEAD, starts the all-reduce of everything but net3DV_1's
        // gradients on a side stream, and issues L1 -- the collective hides under passes C / D
        int stages = (phases & FACL_PHASE_BACKWARD) ? 3 : 0;
        if (phases & FACL_PHASE_BACKWARD_HEAD) stages |= 1;
        else if (phases & FACL_PHASE_BACKWARD_HEAD_X) stages |= 8;
        if (phases & FACL_PHASE_BACKWARD_L1) stages |= 2;
        if (stages && (rc = encoder_backward(d, a->params, a->xt, a->enc_buffers, a->dx, a->dx_global, a->grads, stages, st))) return rc;
    }
    if (phases & FACL_PHASE_UPDATE) {
        if ((rc = facl_adam_step(a->adam_table, a->adam_ntensors, a->lr, a->beta1, a->beta2, a->eps, a->step, stream))) return rc;
        if (a->loss_host) FACL_CHECK(cudaMemcpyAsync(a->loss_host, a->loss2 + 2, sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

}  // extern "C"
