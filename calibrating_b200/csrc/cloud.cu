// cloud.cu -- depth image <-> point cloud (calibrating/utils.py:213-250 depth_to_point_cloud, :254-288 point_cloud_to_arr2d /
// point_cloud_to_depth, :291-317 uvzs_to_arr2d), sm_100a.  SURVEY.md section 8(f) rank 3: the data-parallel neighbours of
// Cam.project_cam2_depth, offered stand-alone.
//
// depth_to_point_cloud keeps NumPy's order (the valid pixels of the -- optionally nearest-neighbour up-sampled -- depth image in
// row-major order), so the compaction is an ordered one: per-row counts, an exclusive scan over the rows, then every warp packs
// its row with ballots.  point_cloud_to_depth is a z-buffered splat: the reference sorts by descending z and lets the last
// write win, i.e. the smallest z per pixel survives = atomicMin on an order-preserving 64-bit key of z.
// Both are HBM-bound (24-40 B per point); float64 like the reference.
#include "b2s_internal.h"

namespace {

struct CloudArgs {
    int W, H, Wu, Hu; // source and up-sampled size
    double rate, sx, sy;
    double Kinv[9];
};

// cv2.resize INTER_NEAREST: source index = min(floor(dst * scale), size - 1)
__device__ __forceinline__ double sample(const double *__restrict__ depth, const CloudArgs &p, int uu, int vv)
{
    const int su = min((int)floor(uu * p.sx), p.W - 1), sv = min((int)floor(vv * p.sy), p.H - 1);
    return depth[(size_t)sv * p.W + su];
}

// block = one up-sampled row: number of non-zero samples
__global__ void __launch_bounds__(256) cloud_count_kernel(const double *__restrict__ depth, CloudArgs p, unsigned *__restrict__ rowcount)
{
    const int vv = blockIdx.x;
    unsigned n = 0;
    for (int uu = threadIdx.x; uu < p.Wu; uu += blockDim.x) n += sample(depth, p, uu, vv) != 0.0;
    n = __reduce_add_sync(0xffffffffu, n);
    __shared__ unsigned part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += part[i];
        rowcount[vv] = t;
    }
}

// one block: exclusive scan of the row counts (rows <= a few thousand), total in rowoff[Hu]
__global__ void __launch_bounds__(1024) cloud_scan_kernel(const unsigned *__restrict__ rowcount, unsigned long long *__restrict__ rowoff, int Hu)
{
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < Hu; base += 1024) {
        const int i = base + threadIdx.x;
        unsigned long long v = i < Hu ? rowcount[i] : 0, incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = wsum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if ((int)threadIdx.x >= o) wi += t;
            }
            wsum[threadIdx.x] = wi - w; // exclusive
        }
        __syncthreads();
        if (i < Hu) rowoff[i] = carry + wsum[threadIdx.x >> 5] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += wsum[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) rowoff[Hu] = carry;
}

// warp = one up-sampled row: ordered compaction with ballots, point = Kinv @ (u z, v z, z), u = column / rate
template <int COLS>
__global__ void __launch_bounds__(256) cloud_emit_kernel(const double *__restrict__ depth, CloudArgs p, const unsigned long long *__restrict__ rowoff,
                                                         double *__restrict__ out, unsigned long long capacity)
{
    const int vv = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (vv >= p.Hu) return;
    unsigned long long o = rowoff[vv];
    const double v = (double)vv / p.rate;
    for (int base = 0; base < p.Wu; base += 32) {
        const int uu = base + lane;
        const double z = uu < p.Wu ? sample(depth, p, uu, vv) : 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, z != 0.0);
        if (z != 0.0) {
            const unsigned long long k = o + __popc(m & ((1u << lane) - 1));
            if (k < capacity) {
                const double u = (double)uu / p.rate, a0 = u * z, a1 = v * z;
                double *q = out + k * COLS;
                q[0] = p.Kinv[0] * a0 + p.Kinv[1] * a1 + p.Kinv[2] * z;
                q[1] = p.Kinv[3] * a0 + p.Kinv[4] * a1 + p.Kinv[5] * z;
                q[2] = p.Kinv[6] * a0 + p.Kinv[7] * a1 + p.Kinv[8] * z;
                if (COLS == 5) {
                    q[3] = u;
                    q[4] = v;
                }
            }
        }
        o += __popc(m);
    }
}

__device__ __forceinline__ unsigned long long zkey(double z)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(z);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull); // order-preserving: smaller z <-> smaller key
}
__device__ __forceinline__ double zunkey(unsigned long long k)
{
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
}

// thread = one point: xyz @ K^T, /z, round half to even (np.round), bounds check, smallest z per pixel wins
__global__ void __launch_bounds__(256) cloud_splat_kernel(const double *__restrict__ pts, unsigned long long n, const double *__restrict__ K, int W, int H,
                                                          unsigned long long *__restrict__ key)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X = pts[i * 3], Y = pts[i * 3 + 1], Z = pts[i * 3 + 2];
    const double px = K[0] * X + K[1] * Y + K[2] * Z, py = K[3] * X + K[4] * Y + K[5] * Z, pz = K[6] * X + K[7] * Y + K[8] * Z;
    const double fu = rint(px / pz), fv = rint(py / pz);
    if (!(fu >= 0.0 && fu < (double)W && fv >= 0.0 && fv < (double)H)) return; // (NaN and Inf fail the comparisons, like the int32 cast + mask)
    atomicMin(&key[(size_t)(int)fv * W + (int)fu], zkey(pz));
}
__global__ void cloud_resolve_kernel(const unsigned long long *__restrict__ key, double *__restrict__ out, size_t n, double bg)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    out[i] = k == 0xFFFFFFFFFFFFFFFFull ? bg : zunkey(k);
}

// ---- calibrating/utils.py:347-415 interpolate_uvzs, the dense half: one thread per pixel of the (h, w) output ----
// "lstsq": z = a u + b v + c on the pixels of the mask (float64 like NumPy's promotion of float32 @ float64, stored float32)
__global__ void plane_fill_kernel(const uint8_t *__restrict__ mask, int H, int W, double a, double b, double c, float *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    // unfused, left to right: the float64 evaluation order of the reference's `output_uvzs @ abc` as numpy runs it
    out[i] = (!mask || mask[i]) ? (float)__dadd_rn(__dadd_rn(__dmul_rn((double)x, a), __dmul_rn((double)y, b)), c) : 0.f;
}
// "nearest": the z of the nearest sparse point if it is closer than `distance`, else 0 (KDTree.query + the distance test);
// brute force over the points, staged through shared memory in tiles (a few hundred to a few thousand points)
__global__ void __launch_bounds__(256) nearest_fill_kernel(const double *__restrict__ uvz, int n, const uint8_t *__restrict__ mask, int H, int W,
                                                           double distance, float *__restrict__ out)
{
    __shared__ double su[256], sv[256], sz[256];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const double px = (double)x, py = (double)y;
    double best = 1e300, bz = 0.0;
    for (int base = 0; base < n; base += 256) {
        const int k = base + threadIdx.x;
        if (k < n) {
            su[threadIdx.x] = uvz[3 * k];
            sv[threadIdx.x] = uvz[3 * k + 1];
            sz[threadIdx.x] = uvz[3 * k + 2];
        }
        __syncthreads();
        const int m = min(256, n - base);
        for (int j = 0; j < m; j++) {
            const double du = su[j] - px, dv = sv[j] - py, d2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));
            if (d2 < best) { // (strict: of equally near points the first one wins)
                best = d2;
                bz = sz[j];
            }
        }
        __syncthreads();
    }
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    out[i] = ((!mask || mask[i]) && sqrt(best) < distance) ? (float)bz : 0.f;
}

// "rbf" (thin plate): sum_i w_i * r^2 * log(r) = sum_i w_i * 0.5 * r2 * log(r2); float64, samples staged through shared memory
__global__ void __launch_bounds__(256) rbf_fill_kernel(const double *__restrict__ uvw, int n, const uint8_t *__restrict__ mask, int H, int W,
                                                       double *__restrict__ out)
{
    __shared__ double su[256], sv[256], sw[256];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const double px = (double)x, py = (double)y;
    double acc = 0.0;
    for (int base = 0; base < n; base += 256) {
        const int k = base + threadIdx.x;
        if (k < n) {
            su[threadIdx.x] = uvw[3 * k];
            sv[threadIdx.x] = uvw[3 * k + 1];
            sw[threadIdx.x] = uvw[3 * k + 2];
        }
        __syncthreads();
        const int m = min(256, n - base);
        for (int j = 0; j < m; j++) {
            const double du = su[j] - px, dv = sv[j] - py, r2 = du * du + dv * dv;
            if (r2 > 0.0) acc += sw[j] * (0.5 * r2 * log(r2));
        }
        __syncthreads();
    }
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    out[i] = (!mask || mask[i]) ? acc : 0.0;
}

} // namespace

cudaError_t launch_rbf_fill(b2s_ctx *c, const double *d_uvw, int n, const uint8_t *d_mask, int H, int W, double *d_out)
{
    rbf_fill_kernel<<<dim3((W + 255) / 256, H), 256, 0, c->stream>>>(d_uvw, n, d_mask, H, W, d_out);
    c->launches++;
    return cudaGetLastError();
}
cudaError_t launch_plane_fill(b2s_ctx *c, const uint8_t *d_mask, int H, int W, double a, double b, double cc, float *d_out)
{
    plane_fill_kernel<<<dim3((W + 255) / 256, H), 256, 0, c->stream>>>(d_mask, H, W, a, b, cc, d_out);
    c->launches++;
    return cudaGetLastError();
}
cudaError_t launch_nearest_fill(b2s_ctx *c, const double *d_uvz, int n, const uint8_t *d_mask, int H, int W, double distance, float *d_out)
{
    nearest_fill_kernel<<<dim3((W + 255) / 256, H), 256, 0, c->stream>>>(d_uvz, n, d_mask, H, W, distance, d_out);
    c->launches++;
    return cudaGetLastError();
}

// d_depth (H, W) f64 -> d_out (n, cols) f64 (cols = 3, or 5 with u, v appended), n written to d_rowoff[Hu] (device); Hu returned
cudaError_t launch_depth_to_cloud(b2s_ctx *c, const double *d_depth, int H, int W, double rate, const double *Kinv, int cols, double *d_out,
                                  unsigned long long capacity, unsigned *d_rowcount, unsigned long long *d_rowoff, int *Hu_out)
{
    CloudArgs p;
    p.W = W; p.H = H; p.rate = rate;
    p.Wu = rate == 1.0 ? W : (int)nearbyint(W * rate); // int(round(x * rate)), Python's round-half-even
    p.Hu = rate == 1.0 ? H : (int)nearbyint(H * rate);
    if (p.Wu <= 0 || p.Hu <= 0) return cudaErrorInvalidValue;
    p.sx = 1.0 / ((double)p.Wu / (double)W); // cv::resize: ifx = 1 / inv_scale_x, inv_scale_x = dsize.width / ssize.width
    p.sy = 1.0 / ((double)p.Hu / (double)H);
    for (int i = 0; i < 9; i++) p.Kinv[i] = Kinv[i];
    cloud_count_kernel<<<p.Hu, 256, 0, c->stream>>>(d_depth, p, d_rowcount);
    cloud_scan_kernel<<<1, 1024, 0, c->stream>>>(d_rowcount, d_rowoff, p.Hu);
    if (cols == 5) cloud_emit_kernel<5><<<(p.Hu + 7) / 8, 256, 0, c->stream>>>(d_depth, p, d_rowoff, d_out, capacity);
    else cloud_emit_kernel<3><<<(p.Hu + 7) / 8, 256, 0, c->stream>>>(d_depth, p, d_rowoff, d_out, capacity);
    c->launches += 3;
    *Hu_out = p.Hu;
    return cudaGetLastError();
}

cudaError_t launch_cloud_to_depth(b2s_ctx *c, const double *d_pts, unsigned long long n, const double *d_K, int W, int H, unsigned long long *d_key,
                                  double *d_out, double bg)
{
    const size_t npx = (size_t)W * H;
    cudaError_t e = cudaMemsetAsync(d_key, 0xFF, npx * 8, c->stream);
    if (e != cudaSuccess) return e;
    if (n) cloud_splat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_pts, n, d_K, W, H, d_key);
    cloud_resolve_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, c->stream>>>(d_key, d_out, npx, bg);
    c->launches += n ? 2 : 1;
    return cudaGetLastError();
}
