// sgbm_wave.cu -- semi-global path aggregation as two dependency-ordered wavefront sweeps, sm_100a.
//
// Replaces the aggregation loops of cv2.StereoSGBM.compute (calibrating/stereo_matching.py:63) in
// OpenCV's own order: pass 1 carries the four paths (+1,0) (+1,+1) (0,+1) (-1,+1) top-down, pass 2 the four mirrored paths
// bottom-up; each pass reads C once and writes the saturated sum of its four L_r once -- the canonical 2 + 2 bytes per voxel
// and sweep of SURVEY.md section 8(d).  (The round-1 schedule of sgbm_agg.cu splits a pass into horizontal scans and a
// lock-step vertical sweep and moves 20 B/voxel.)
//
//   * Logical coordinates (u, v): top-down sweep u = x1, v = y; bottom-up sweep u = width1-1-x1, v = H-1-y.  In (u, v) both
//     sweeps are the same program: pixel (u, v) needs  path h: (u-1, v)   dr: (u-1, v-1)   d: (u, v-1)   dl: (u+1, v-1).
//     A row runs two pixels behind the row above it and never waits for a row below: dependencies point one way, nothing
//     is lock-stepped, a late row only delays its successors.
//   * One WARP owns one image row for the whole launch, and its four groups of 8 lanes own the four PATHS of the current
//     pixel: one SGM step (packed int16x2 DPX arithmetic, the +-1 disparity neighbours by an 8-wide shuffle, the minimum by
//     a 3-level butterfly) advances all four paths at once.  A lane holds 4*NP packed registers = 8*NP disparities of one
//     path; the cost volume uses the block layout (SgbmGeom::layout 1: a 32-bit word = disparities (16b+j, 16b+8+j)), so the
//     d-1 / d+1 neighbours of a word are whole registers and only one word per block of 8 needs a PRMT.
//   * Path h keeps its state in registers; the three row-crossing paths take the state of the row above from that row's
//     ring in shared memory (8 pixels x {dr, d, dl}; each group reads ITS part of ITS pixel: u-1, u, u+1) and publish their
//     own for the row below.  A full-mbarrier per ring slot orders producer and consumer warp, a progress counter keeps the
//     producer from overrunning the ring.  The four L_r are summed through a small exchange buffer and stored once.
//   * C arrives by 1-D bulk copies (cp.async.bulk -> UBLKCP, mbarrier per chunk of 4 pixels).
//   * Bands of 16 rows (one CTA) are chained through a ring in global memory whose int16 words carry a phase bit (states are
//     normalised, L - minL < 2^15, so bit 15 is free): no flags, no fences.  The consuming band copies slots into shared
//     memory with cp.async six pixels ahead of their use and validates the phase bits of the words it reads; a miss
//     re-synchronises once (wait for the far end of the window, refill).  Band numbers come from an atomic ticket, so a
//     band's predecessor is always resident or finished: no co-residency requirement, no cooperative launch, no ordering
//     between launches of different handles.
//   * Both sweeps of MODE_HH share ONE launch (tickets alternate between them) and write separate sums S (top-down) and S2
//     (bottom-up); wta_kernel (sgbm_post.cu) adds them.
#include <stdlib.h>

#include <mutex>

#include "sgm_common.cuh"

namespace {

constexpr int WV_K = 8;      // pixels per state ring (power of two)
constexpr int WV_CPX = 4;    // pixels per bulk-copy chunk
constexpr int WV_KG = 32;    // pixels per band hand-over ring in global memory (power of two)
constexpr int WV_KS = 16;    // pixels of the hand-over staging ring in shared memory (power of two)
constexpr int WV_L2AHEAD = 24; // chunks of C pulled into L2 ahead of the copy into shared memory
constexpr int WV_PF = 6;     // prefetch distance of the hand-over (pixels; < WV_KS - 3)
template <int NP> struct WvCfg {
    static constexpr int rows = NP <= 2 ? 16 : 8;   // image rows (= warps) per CTA
    static constexpr int cslots = NP <= 2 ? 3 : 2;  // chunk slots per row
};

struct WaveArgs {
    const int16_t *C;
    int16_t *S, *S2; // sums of the top-down / bottom-up sweep
    int H, width1, D, P1, P2;
    int ndirs, dir0; // sweeps in this launch (2: tickets alternate top-down / bottom-up; 1: only dir0)
    int nbands;
    uint32_t *gring; // [2][nbands][WV_KG][3][32*NP] u32, memset to 0xFF before the launch (phase 1)
    int *gcons;      // [2][nbands] pixels the band has taken from its predecessor's ring, zeroed before the launch
    int *ticket;     // zeroed before the launch
    int *err;
};

template <int N> __device__ __forceinline__ void lds_n(uint32_t addr, uint32_t (&v)[N])
{
    static_assert(N % 4 == 0, "whole 16-byte vectors");
#pragma unroll
    for (int i = 0; i < N; i += 4)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i]), "=r"(v[i + 1]), "=r"(v[i + 2]), "=r"(v[i + 3]) : "r"(addr + 4 * i) : "memory");
}
template <int N> __device__ __forceinline__ void sts_n(uint32_t addr, const uint32_t (&v)[N])
{
#pragma unroll
    for (int i = 0; i < N; i += 4)
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr + 4 * i), "r"(v[i]), "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3]) : "memory");
}

template <int NP, bool PAD>
__global__ void __launch_bounds__(WvCfg<NP>::rows * 32, 1) agg_wave_kernel(WaveArgs a)
{
    constexpr int ROWS = WvCfg<NP>::rows, CSLOTS = WvCfg<NP>::cslots;
    constexpr int N = 4 * NP;             // packed registers per lane (one path, 8*NP disparities)
    constexpr int CH = 128 * NP;          // bytes of one pixel's d-chunk
    constexpr int DW = 32 * NP;           // ... in 32-bit words
    constexpr int SLOTB = 3 * CH;         // one state-ring slot: parts dr, d, dl
    constexpr int RINGB = (WV_K + 1) * SLOTB; // one row's state ring; slot WV_K holds the out-of-image state for ever
    constexpr int CSLOTB = WV_CPX * CH;   // one chunk slot
    constexpr int CRINGB = CSLOTS * CSLOTB;
    constexpr int LBUFB = 2 * 4 * CH;     // per-row exchange buffer of the four L_r, double-buffered
    constexpr int STGB = WV_KS * SLOTB;   // hand-over staging ring (+ one zero slot)
    constexpr int NCP = (SLOTB / 16 + 31) / 32; // cp.async instructions per lane and hand-over slot
    extern __shared__ __align__(128) unsigned char wv_smem[];
    // shared memory: C rings [ROWS] | state rings [ROWS] | L buffers [ROWS] | staging ring + zero slot | chunk mbarriers
    // [ROWS][CSLOTS] | full mbarriers [ROWS][K] | consumer progress [ROWS] | ticket
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(wv_smem);
    const uint32_t sr0 = sm0 + ROWS * CRINGB;
    const uint32_t lb0 = sr0 + ROWS * RINGB;
    const uint32_t sg0 = lb0 + ROWS * LBUFB;
    const uint32_t cb0 = sg0 + STGB + SLOTB;
    const uint32_t fb0 = cb0 + ROWS * CSLOTS * 8;
    volatile int *cons = (volatile int *)(wv_smem + (fb0 - sm0) + ROWS * WV_K * 8);
    int *tick = (int *)(cons + ROWS);

    if (threadIdx.x == 0) *tick = atomicAdd(a.ticket, 1);
    if (threadIdx.x < ROWS * CSLOTS) mbar_init(cb0 + threadIdx.x * 8, 1);
    if (threadIdx.x < ROWS * WV_K) mbar_init(fb0 + threadIdx.x * 8, 1);
    if (threadIdx.x < ROWS) cons[threadIdx.x] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int t = *tick;
    const int dir = a.ndirs == 2 ? (t & 1) : a.dir0;
    const int band = a.ndirs == 2 ? (t >> 1) : t;

    const int lane = threadIdx.x & 31;
    const int r = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); // row inside the band = warp
    const int g = lane >> 3, li = lane & 7;                             // path (0 h, 1 dr, 2 d, 3 dl), lane inside the group
    const int v = band * ROWS + r;                                      // logical row
    const int H = a.H, W = a.width1, Dp = 64 * NP;
    if (v >= H) return; // (rows above do not wait for rows that do not exist)
    const int y = dir ? H - 1 - v : v;
    const uint32_t BIG = 0x7FFF7FFFu;

    uint32_t padmask[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int wd = li * N + i, d0 = (wd >> 3) * 16 + (wd & 7);
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d0 + 8 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;
    uint32_t ku = li != 0 ? 1u : 0u, kd = li != 7 ? 1u : 0u;
    asm("" : "+r"(ku)); // (keeps the compiler from turning the multiply-adds back into selects)
    asm("" : "+r"(kd));
    const uint32_t au = li == 0 ? BIG : 0u, ad = li == 7 ? BIG : 0u;

    // the out-of-image state (L = 0, minL = 0) in slot K of this row's ring and, for the first row of a band, in the zero slot
    // behind the staging ring
    const uint32_t myring = sr0 + r * RINGB;
    if (g > 0) sts_n<N>(myring + WV_K * SLOTB + (g - 1) * CH + li * N * 4, padmask);
    if (r == 0 && g > 0) sts_n<N>(sg0 + STGB + (g - 1) * CH + li * N * 4, padmask);
    __syncwarp();

    const int16_t *Crow = a.C + (size_t)y * W * Dp;
    int16_t *Orow = (dir ? a.S2 : a.S) + (size_t)y * W * Dp + lane * 2 * NP;
    const uint32_t cring = sm0 + r * CRINGB, cbar = cb0 + r * CSLOTS * 8;
    const int nchunks = (W + WV_CPX - 1) / WV_CPX;
    auto issue = [&](int j) { // chunk j = pixels [CPX*j, CPX*j + CPX) of this row
        if (lane == 0) {
            const int ua = WV_CPX * j, ub = min(ua + WV_CPX, W);
            const uint32_t slot = cring + (j % CSLOTS) * CSLOTB, mb = cbar + (j % CSLOTS) * 8;
            const uint32_t bytes = (uint32_t)(ub - ua) * CH;
            mbar_arrive_expect_tx(mb, bytes);
            if (dir == 0) bulk_g2s(slot, Crow + (size_t)ua * Dp, bytes, mb);
            else bulk_g2s(slot + (WV_CPX - (ub - ua)) * CH, Crow + (size_t)(W - ub) * Dp, bytes, mb); // memory order = reversed pixel order
        }
    };
    // DRAM latency (~2 us under load) against a prefetch distance of (CSLOTS - 1) * CPX = 8 pixels in shared memory would hold a
    // row at ~500 cycles per pixel: chunks are pulled into L2 WV_L2AHEAD chunks ahead (cp.async.bulk.prefetch.L2), so that the
    // copy into shared memory only pays the L2 latency
    auto l2_prefetch = [&](int j) {
        if (lane == 0 && j < nchunks) {
            const int ua = WV_CPX * j, ub = min(ua + WV_CPX, W);
            const int16_t *src = dir == 0 ? Crow + (size_t)ua * Dp : Crow + (size_t)(W - ub) * Dp;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(ub - ua) * CH) : "memory");
        }
    };
#pragma unroll 1
    for (int j = 0; j < WV_L2AHEAD; j++) l2_prefetch(j);
#pragma unroll 1
    for (int j = 0; j < CSLOTS - 1 && j < nchunks; j++) issue(j);

    // ---- where this lane's group reads its predecessor state ----
    // g = 1, 2, 3: part g-1 of pixel u-1, u, u+1 of the row above (the ring of row r-1, or the staging ring for the first row of
    // a band, whose words still carry the phase bit); out-of-image predecessors read the zero slot.
    const bool first_band = band == 0, takes = r == 0 && !first_band;
    const int off = g == 3 ? 1 : (g == 2 ? 0 : -1);
    const uint32_t part = (uint32_t)(g > 0 ? g - 1 : 0) * CH + li * N * 4;
    const uint32_t upbase = (r == 0 ? sg0 : sr0 + (r - 1) * RINGB) + part;
    const uint32_t upzero = (r == 0 ? sg0 + STGB : sr0 + (r - 1) * RINGB + WV_K * SLOTB) + part;
    const int upmask = r == 0 ? WV_KS - 1 : WV_K - 1;
    const bool zalways = v == 0; // the first logical row: every predecessor is outside the image
    const uint32_t pubbase = myring + part;

    // ---- hand-over between bands ----
    const bool gives = band + 1 < a.nbands && r == ROWS - 1; // this row's states go to the next band
    const size_t gslotw = (size_t)3 * DW;
    const uint32_t *gin = a.gring + ((size_t)dir * a.nbands + (first_band ? 0 : band - 1)) * WV_KG * gslotw;
    uint32_t *gout = a.gring + ((size_t)dir * a.nbands + band) * WV_KG * gslotw + (size_t)(g > 0 ? g - 1 : 0) * DW + li * N;
    volatile int *gcons_mine = a.gcons + dir * a.nbands + band;                        // written by this band (as consumer)
    volatile int *gcons_next = a.gcons + dir * a.nbands + min(band + 1, a.nbands - 1); // read by this band (as producer)
    auto stage = [&](int p) { // asynchronous copy of pixel p's hand-over slot into the staging ring
        if (p < W) {
            const unsigned char *src = (const unsigned char *)(gin + (size_t)(p & (WV_KG - 1)) * gslotw);
            const uint32_t dst = sg0 + (p & (WV_KS - 1)) * SLOTB;
#pragma unroll
            for (int k = 0; k < NCP; k++) {
                const int seg = lane + 32 * k;
                if (seg < SLOTB / 16) cp_async16_s(dst + seg * 16, src + seg * 16);
            }
        }
        cp_async_commit();
    };
    // expected phase word of the pixel this group reads in step u
    auto phase_of = [&](int p) -> uint32_t { return ((p / WV_KG) & 1) ? 0x80008000u : 0u; };
    // a staged word had the wrong phase: the producer is not WV_PF pixels ahead.  Wait until it is (poll the far end of the
    // window in global memory), then refill the window.
    auto resync = [&](int s) {
        const int far = min(s + 1 + WV_PF, W - 1);
        const uint32_t *src = gin + (size_t)(far & (WV_KG - 1)) * gslotw + lane;
        const uint32_t ph = phase_of(far);
        int spins = 0;
        unsigned long long t0 = 0;
        while (true) {
            uint32_t bad = 0;
#pragma unroll
            for (int k = 0; k < 3 * NP; k++) {
                uint32_t wv;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(wv) : "l"(src + 32 * k) : "memory");
                bad |= (wv ^ ph) & 0x80008000u;
            }
            if (__all_sync(0xffffffffu, bad == 0)) break;
            if (wait_expired(++spins, t0, a.err)) {
                *(volatile int *)a.err = 1;
                break;
            }
        }
        cp_async_wait<0>();
        __syncwarp();
        for (int p = max(s - 1, 0); p <= far; p++) stage(p);
        cp_async_wait<0>();
        __syncwarp();
    };
    if (takes) {
#pragma unroll 1
        for (int p = 0; p <= WV_PF; p++) stage(p); // pixels 0 .. WV_PF: one group each
    }
    int cons_seen = 0; // producer side: last value read from the consumer's progress counter

    uint32_t T[N];
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = padmask[i];
    const bool waits = r > 0; // the row above lives in this CTA: wait on its full-mbarriers
    const uint32_t fb_up = fb0 + (r - 1) * WV_K * 8, fb_my = fb0 + r * WV_K * 8;
    const bool has_next = r + 1 < ROWS && v + 1 < H; // a row of this CTA consumes this row's ring
    if (waits) mbar_wait_sleep(fb_up, 0, a.err); // pixel 0 of the row above

    int s = 0;
#pragma unroll 1
    for (int j = 0; j < nchunks; j++) {
        __syncwarp(); // every lane is done with the slot of chunk j-1: it is refilled now
        if (j + CSLOTS - 1 < nchunks) issue(j + CSLOTS - 1);
        l2_prefetch(j + WV_L2AHEAD);
        mbar_wait_sleep(cbar + (j % CSLOTS) * 8, (uint32_t)(j / CSLOTS) & 1u, a.err);
        const int np = min(WV_CPX, W - j * WV_CPX);
        const uint32_t cslot = cring + (j % CSLOTS) * CSLOTB + li * N * 4;
#pragma unroll 1
        for (int p = 0; p < np; p++, s++) {
            uint32_t c[N], L[N];
            lds_n<N>(cslot + (dir ? (WV_CPX - 1 - p) : p) * CH, c);
            // ---- the row above: pixel s+1 of it must be there ----
            if (takes) {
                stage(s + 1 + WV_PF);
                cp_async_wait<WV_PF>(); // the group of pixel s+1 has landed
                __syncwarp();
            } else if (waits && s + 1 < W)
                mbar_wait_sleep(fb_up + ((s + 1) & (WV_K - 1)) * 8, (uint32_t)((s + 1) / WV_K) & 1u, a.err);
            const int pu = s + off; // pixel of the row above whose state this group needs
            const bool zero = zalways || pu < 0 || pu >= W;
            const uint32_t upaddr = zero ? upzero : upbase + (pu & upmask) * SLOTB;
            if (g > 0) lds_n<N>(upaddr, T);
            if (takes) {
                // the staged words still carry the phase bit of the pixel they were written for: a mismatch means the copy ran
                // ahead of the producer
                const uint32_t ph = (zero || g == 0) ? 0u : phase_of(pu);
                while (true) {
                    uint32_t bad = 0;
                    if (g > 0) {
#pragma unroll
                        for (int i = 0; i < N; i++) bad |= (T[i] ^ ph) & 0x80008000u;
                    }
                    if (__all_sync(0xffffffffu, bad == 0)) break;
                    if (*(volatile int *)a.err != 0) break;
                    resync(s);
                    if (g > 0) lds_n<N>(upaddr, T);
                }
#pragma unroll
                for (int i = 0; i < N; i++) T[i] &= 0x7FFF7FFFu;
                if (lane == 0 && (s & 7) == 7) *gcons_mine = s; // pixels <= s-1 of the predecessor's ring are no longer needed
            }
            // ---- all four paths of pixel s ----
            sgm_step_blk<N, 8, PAD>(T, c, L, padmask, P1v, P2mP1v, ku, au, kd, ad);
            // ---- publish the row-crossing states of pixel s ----
            if (gives) {
                const int need = s - WV_KG + 9; // the slot held pixel s-KG; the consumer reports its progress every 8 pixels
                if (cons_seen < need) {
                    int spins = 0;
                    unsigned long long t0 = 0;
                    while ((cons_seen = *gcons_next) < need) {
                        if (wait_expired(++spins, t0, a.err)) {
                            *(volatile int *)a.err = 1;
                            break;
                        }
                    }
                }
                if (g > 0) {
                    const uint32_t ph = phase_of(s);
                    uint32_t *dst = gout + (size_t)(s & (WV_KG - 1)) * gslotw;
#pragma unroll
                    for (int i = 0; i < N; i += 4)
                        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"((T[i] & 0x7FFF7FFFu) | ph),
                                     "r"((T[i + 1] & 0x7FFF7FFFu) | ph), "r"((T[i + 2] & 0x7FFF7FFFu) | ph), "r"((T[i + 3] & 0x7FFF7FFFu) | ph)
                                     : "memory");
                }
            } else if (has_next) {
                const int need = s - WV_K + 2; // the slot held pixel s-K; the consumer must be past it
                if (cons_seen < need) {
                    int spins = 0;
                    unsigned long long t0 = 0;
                    while ((cons_seen = cons[r + 1]) < need) {
                        if (wait_expired(++spins, t0, a.err)) {
                            *(volatile int *)a.err = 1;
                            break;
                        }
                    }
                }
                if (g > 0) sts_n<N>(pubbase + (s & (WV_K - 1)) * SLOTB, T);
            }
            // ---- sum of the four L_r through the exchange buffer, one coalesced store ----
            const uint32_t lb = lb0 + r * LBUFB + (s & 1) * (4 * CH);
            sts_n<N>(lb + g * CH + li * N * 4, L);
            __syncwarp(); // states and L_r of this step are in shared memory
            if (has_next && lane == 0) mbar_arrive(fb_my + (s & (WV_K - 1)) * 8);
            uint32_t acc[NP], q[NP];
            lds_s<NP>(lb + lane * NP * 4, acc);
#pragma unroll
            for (int k = 1; k < 4; k++) {
                lds_s<NP>(lb + k * CH + lane * NP * 4, q);
#pragma unroll
                for (int i = 0; i < NP; i++) acc[i] = __viaddmin_u16x2(acc[i], q[i], BIG); // saturating sums (L >= 0: any order)
            }
            stcg_regs<NP>(Orow + (size_t)(dir ? W - 1 - s : s) * Dp, acc);
            if (lane == 0) cons[r] = s + 1; // this row is done with pixels <= s-1 of the row above
        }
    }
}

template <int NP, bool PAD> cudaError_t launch_wave_t(b2s_ctx *c, const WaveArgs &a)
{
    constexpr int ROWS = WvCfg<NP>::rows, CSLOTS = WvCfg<NP>::cslots, CH = 128 * NP;
    const size_t smem = (size_t)ROWS * CSLOTS * WV_CPX * CH + (size_t)ROWS * (WV_K + 1) * 3 * CH + (size_t)ROWS * 8 * CH + (size_t)(WV_KS + 1) * 3 * CH +
                        ROWS * CSLOTS * 8 + ROWS * WV_K * 8 + ROWS * 4 + 16;
    static std::once_flag once[64]; // per instantiation and device (the attribute belongs to the device's context)
    cudaError_t e = cudaSuccess;
    std::call_once(once[c->device & 63], [&] { e = cudaFuncSetAttribute(agg_wave_kernel<NP, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    if (e != cudaSuccess) return e;
    agg_wave_kernel<NP, PAD><<<a.ndirs * a.nbands, ROWS * 32, smem, c->stream>>>(a);
    c->launches++;
    return cudaGetLastError();
}

} // namespace

int wave_rows_per_band(int NP) { return NP <= 2 ? WvCfg<2>::rows : WvCfg<4>::rows; }

// ndirs = 2: both sweeps of MODE_HH (S = top-down sum, S2 = bottom-up sum); ndirs = 1: the top-down sweep alone (S).
// Needs the block layout of the cost volume (SgbmGeom::layout 1, NP = 2 or 4).
cudaError_t launch_wave(b2s_ctx *c, int ndirs)
{
    const SgbmGeom &g = c->g;
    if (g.layout != 1 || (g.NP != 2 && g.NP != 4)) return cudaErrorInvalidValue;
    if (cudaError_t te = wait_timeout_init()) return te;
    WaveArgs a;
    a.C = c->C.as<int16_t>();
    a.S = c->S.as<int16_t>();
    a.S2 = c->S2.as<int16_t>();
    a.H = g.H; a.width1 = g.width1; a.D = g.D; a.P1 = g.P1; a.P2 = g.P2;
    a.ndirs = ndirs; a.dir0 = 0;
    const int rows = wave_rows_per_band(g.NP);
    a.nbands = (g.H + rows - 1) / rows;
    const size_t ring_bytes = (size_t)2 * a.nbands * WV_KG * 3 * 32 * g.NP * sizeof(uint32_t);
    const size_t ctl_bytes = ((size_t)2 * a.nbands + 1) * sizeof(int);
    cudaError_t e = c->agg_ho.ensure(ring_bytes + ctl_bytes);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->agg_ho.p, 0xFF, ring_bytes, c->stream)) != cudaSuccess) return e; // phase 1 everywhere
    if ((e = cudaMemsetAsync((char *)c->agg_ho.p + ring_bytes, 0, ctl_bytes, c->stream)) != cudaSuccess) return e;
    a.gring = c->agg_ho.as<uint32_t>();
    a.gcons = (int *)((char *)c->agg_ho.p + ring_bytes);
    a.ticket = a.gcons + 2 * a.nbands;
    a.err = c->agg_err;
    const bool pad = g.D != g.Dp;
    if (g.NP == 2) return pad ? launch_wave_t<2, true>(c, a) : launch_wave_t<2, false>(c, a);
    return pad ? launch_wave_t<4, true>(c, a) : launch_wave_t<4, false>(c, a);
}
