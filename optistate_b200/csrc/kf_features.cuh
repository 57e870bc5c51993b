// Next row after the filter: the driver's feature assembly and the GRU's input preparation,
// kept on the device so that KF estimates feed the learned corrector without a host round trip.
//
//   assemble   data_conversion_Kalman_to_Training.py:245-254 - one 60-wide row per (trajectory, step):
//              [x(12) | imu_acc(6) | f(12) | p_world(12) | dp(12) | imu(6)]          -> rows [N][T][60]
//   min-max    gru/gru_train.py:56-63   - per-column minimum / maximum over all rows
//   windows    gru/gru_train.py:108-111,180-192 - (v - min) / (max - min) in the filter's precision, the encoder
//              latent appended, sliding windows of `seq_len` rows, cast to float32   -> [G][R-seq+1][seq][60+latent]
// All three are pure data movement: HBM-bound, coalesced through shared-memory transposes / row-contiguous copies.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace okf {

constexpr int FEAT = 60;

// block = 64 trajectories x FEAT_STEPS steps; 256 threads.  Loads are coalesced along the trajectory index (SoA
// inputs, 64 consecutive values per channel), stores are row-contiguous (60 values per trajectory row).
constexpr int FEAT_TRAJ = 64, FEAT_STEPS = 4;

template <typename Real>
__global__ void __launch_bounds__(256) kf_features_kernel(long long N, long long T, long long S, long long stream_offset,
                                                           const int32_t *__restrict__ stream_index, const Real *__restrict__ x_steps,
                                                           const Real *__restrict__ p_world, const Real *__restrict__ imu,
                                                           const Real *__restrict__ imu_acc, const Real *__restrict__ f,
                                                           const Real *__restrict__ dp, Real *__restrict__ rows) {
    __shared__ Real tile[FEAT][FEAT_TRAJ + 1];
    const long long i0 = (long long)blockIdx.x * FEAT_TRAJ;
    const int li_ld = threadIdx.x % FEAT_TRAJ, c_ld0 = threadIdx.x / FEAT_TRAJ;  // 4 channels in flight per pass
    const long long i_ld = i0 + li_ld;
    const long long s_ld = i_ld < N ? (stream_index ? (long long)stream_index[i_ld] : (i_ld + stream_offset) % S) : 0;
    for (long long t = (long long)blockIdx.y * FEAT_STEPS; t < T && t < ((long long)blockIdx.y + 1) * FEAT_STEPS; ++t) {
#pragma unroll 5
        for (int c = c_ld0; c < FEAT; c += 256 / FEAT_TRAJ) {
            Real v = Real(0);
            if (i_ld < N) {
                if (c < 12) v = x_steps[(t * 12 + c) * N + i_ld];
                else if (c < 18) v = imu_acc ? imu_acc[(t * 6 + (c - 12)) * S + s_ld] : Real(0);
                else if (c < 30) v = f[(t * 12 + (c - 18)) * S + s_ld];
                else if (c < 42) v = p_world[(t * 12 + (c - 30)) * N + i_ld];
                else if (c < 54) v = dp[(t * 12 + (c - 42)) * S + s_ld];
                else v = imu[(t * 6 + (c - 54)) * S + s_ld];
            }
            tile[c][li_ld] = v;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < FEAT * FEAT_TRAJ; e += 256) {
            const int li = e / FEAT, c = e % FEAT;
            const long long i = i0 + li;
            if (i < N) __stcs(rows + (i * T + t) * FEAT + c, tile[c][li]);
        }
        __syncthreads();
    }
}

// per-column min / max, pass 1: each block reduces a slab of rows into partial[block][2][cols]
template <typename Real>
__global__ void __launch_bounds__(256) kf_minmax_partial_kernel(const Real *__restrict__ rows, long long n_rows, int cols,
                                                                 Real *__restrict__ partial) {
    extern __shared__ unsigned char sm_raw[];
    Real *smn = reinterpret_cast<Real *>(sm_raw), *smx = smn + blockDim.x;
    const int groups = blockDim.x / cols;  // row groups that fit the block; threads beyond groups*cols idle
    const int g = threadIdx.x / cols, c = threadIdx.x % cols;
    Real mn = Real(INFINITY), mx = Real(-INFINITY);
    if (g < groups) {
        for (long long r = (long long)blockIdx.x * groups + g; r < n_rows; r += (long long)gridDim.x * groups) {
            const Real v = rows[r * cols + c];
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
        }
    }
    smn[threadIdx.x] = mn;
    smx[threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x < cols) {
        for (int k = 1; k < groups; ++k) {
            const Real a = smn[k * cols + threadIdx.x], b = smx[k * cols + threadIdx.x];
            mn = a < mn ? a : mn;
            mx = b > mx ? b : mx;
        }
        partial[((long long)blockIdx.x * 2 + 0) * cols + threadIdx.x] = mn;
        partial[((long long)blockIdx.x * 2 + 1) * cols + threadIdx.x] = mx;
    }
}

template <typename Real>
__global__ void kf_minmax_final_kernel(const Real *__restrict__ partial, int n_partial, int cols, Real *__restrict__ mn_out,
                                       Real *__restrict__ mx_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    Real mn = Real(INFINITY), mx = Real(-INFINITY);
    for (int k = 0; k < n_partial; ++k) {
        const Real a = partial[((long long)k * 2 + 0) * cols + c], b = partial[((long long)k * 2 + 1) * cols + c];
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    mn_out[c] = mn;
    mx_out[c] = mx;
}

// full[r][c] = c < cols ? float((rows[r][c] - mn[c]) / (mx[c] - mn[c])) : latent[r][c - cols]      (each row normalised ONCE)
template <typename Real>
__global__ void __launch_bounds__(256) kf_normalise_rows_kernel(const Real *__restrict__ rows, const float *__restrict__ latent,
                                                                 const Real *__restrict__ mn, const Real *__restrict__ mx, long long n_rows,
                                                                 int cols, int n_latent, float *__restrict__ full) {
    const int width = cols + n_latent;
    const long long total = n_rows * width;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / width;
        const int c = (int)(e - r * width);
        float v;
        if (c < cols) {
            const Real lo = mn[c], hi = mx[c];
            v = (float)((rows[r * cols + c] - lo) / (hi - lo));
        } else {
            v = latent[r * n_latent + (c - cols)];
        }
        full[e] = v;
    }
}

// out[g][n] = full[g*rpg + n .. g*rpg + n + seq) : `seq` consecutive rows are one contiguous run of the normalised buffer,
// so a window is a straight copy of seq*width floats.  Vec = float4 when width % 4 == 0, float otherwise.
template <typename Vec>
__global__ void __launch_bounds__(256) kf_windows_copy_kernel(const Vec *__restrict__ full, long long rows_per_group, long long n_groups,
                                                               int row_vecs, int seq, Vec *__restrict__ out) {
    // blockDim = (64, 4): each thread row copies one window per trip; no per-element index arithmetic
    const long long n_win = rows_per_group - seq + 1;
    const int win_vecs = seq * row_vecs;
    const long long total = n_groups * n_win;
    for (long long w = (long long)blockIdx.x * blockDim.y + threadIdx.y; w < total; w += (long long)gridDim.x * blockDim.y) {
        const long long g = w / n_win, n = w - g * n_win;
        const Vec *src = full + (g * rows_per_group + n) * row_vecs;
        Vec *dst = out + w * win_vecs;
        for (int off = threadIdx.x; off < win_vecs; off += blockDim.x) __stcs(dst + off, src[off]);
    }
}

}  // namespace okf
