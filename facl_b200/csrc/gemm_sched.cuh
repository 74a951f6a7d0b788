// Tile constants and the persistent tile schedule shared by the tcgen05 GEMM kernels (gemm_tc.cu, gemm_img.cu).
#pragma once
#include "gemm_tc.cuh"

namespace facl {
namespace {

constexpr int M_TILE = 128;
constexpr int N_TILE = 256;
constexpr int K_BLK = 64;
constexpr int A_TILE_BYTES = M_TILE * 128;
constexpr int B_TILE_BYTES = N_TILE * 128;
constexpr int NUM_SMS = 148;

struct Work {
    int mt, nt, kb0, kb1;
};

struct Schedule {
    int numMT, numNT, KB, P;
    int mt, nt, step, ks;
    bool split;
    __device__ Schedule(const GemmParams& p) {
        numMT = (p.Md + M_TILE - 1) / M_TILE;
        numNT = (p.Nd + N_TILE - 1) / N_TILE;
        KB = (p.Kd + K_BLK - 1) / K_BLK;
        split = p.ksplit > 1;
        int b = blockIdx.x;
        if (!split) {
            P = gridDim.x / numMT;
            mt = b % numMT;
            nt = b / numMT;   // first n-tile; advance by P
            step = P;
            ks = 0;
        } else {
            P = 1;
            mt = b % numMT;
            nt = (b / numMT) % numNT;
            ks = b / (numMT * numNT);
            step = numNT;     // exactly one item
        }
    }
    __device__ bool get(int it, const GemmParams& p, Work& w) const {
        int n = nt + it * step;
        if (n >= numNT) return false;
        w.mt = mt;
        w.nt = n;
        if (!split) {
            w.kb0 = 0;
            w.kb1 = KB;
        } else {
            w.kb0 = (int)(((long long)ks * KB) / p.ksplit);
            w.kb1 = (int)(((long long)(ks + 1) * KB) / p.ksplit);
        }
        return true;
    }
};

}  // namespace
}  // namespace facl
