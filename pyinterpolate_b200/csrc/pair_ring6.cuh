// pair_ring6_kernel — the ring-window pair kernel, second formulation (included by variogram.cu).
//
// Same job and same exact lag assignment as pair_ring_kernel (reference: distance/point.py:56, :86-92 and
// semivariogram/experimental/functions/general.py:235-303; ellipse: distance/angular.py:641-681), rebuilt around what
// ncu showed for that kernel (profiles/r1_pair_ring_kernel_ncu_full_v5.md): 23.8 warp instructions and 6.7 shared-memory
// wavefronts per warp-wide pair step, shared-memory pipe 73 % busy — the wavefronts alone capped it below 50 % of the
// FP64 peak.  What changed:
//   * z_j FORM.  A ring slot of a thread holds (count, sum z_j, sum z_j^2) of the pairs (i, j) it saw instead of
//     (count, sum dz, sum dz^2): per pair the values cost one select (high word of z_j under the predicate "second lag of
//     the group"), one DADD and one DFMA — no dz.  The fold forms sum dz^2 = c z_i^2 - 2 z_i sum z_j + sum z_j^2,
//     sum z_i z_j = z_i sum z_j, sum (z_i + z_j) = c z_i + sum z_j.  z is centred (mean removed), so the cancellation
//     costs eps * var(z) / (2 gamma(h)) relative; the host side re-runs the dz-form kernel when a lag has
//     gamma(h) < 1e-6 var(z) (variogram_device).
//   * the second-lag predicate can be a 64-bit INTEGER compare of the squared distance's bit pattern (two ALU
//     instructions) instead of a DSETP, which takes that work off the fp64 pipe (PIB_V6_CMP).
//   * ONE 16-byte table entry per lag slot {threshold, high word of the next threshold, both ring offsets} replaces the
//     threshold loads and the ring-slot arithmetic (13 integer instructions per group in the old kernel); the ring size
//     is any number <= 32 (the offsets come from the table, not from a mask).
//   * PER-WARP ring windows: every warp tracks the lag window of ITS 32 * APT A points against the B tile, folds on its
//     own, and skips B tiles none of its points can reach (finer culling than the CTA-wide box).
//   * APT (A points per thread: 1 or 2) and U (pairs per group: 8 or 16) are template parameters: APT = 2 halves the
//     B-point loads per pair, U = 16 halves the ring traffic per pair — the two levers on the shared-memory wavefronts.
//   * optionally (PIB_V6_PIPE) two groups in flight per thread: distances + table lookup of group g + 1 issued ahead of the
//     sums and ring updates of group g.  Measured neutral (the kernel is bound by the shared-memory pipe, not by the
//     latency of the lookup chain), so it is off.
// Measured (B200, BASELINE config 2): 174 ms against 186 ms for pair_ring_kernel; 6.4 shared-memory wavefronts per
// warp-wide pair step (B points 3.25, ring 2.5, tables 0.65) keep the shared-memory pipe at 73 % — 24 warps instead of
// 16, U = 16 (5.2 wavefronts, but 23 % of the groups then span three lags) and APT = 2 (8-12 warps) all land within 5 %
// of the same time; see DESIGN.md section 4.1 for the experiment table.
#pragma once

namespace pib {

struct __align__(16) SlotInfo {
    long long T;        // upper threshold of this slot on the squared distance (bit pattern; d2 >= 0 so integer order = fp order)
    int next_hi;        // high word of the next slot's threshold
    uint32_t offs;      // ring block of this slot (low half) and of the next slot (high half), in 16-byte units
};

template <int RING, int NT, int APT>
static size_t ring6_smem_bytes(int n_slots, int lut_n)
{
    const size_t na = (size_t)NT * APT;
    size_t b = (size_t)(RING + 1) * (na * 16 + (size_t)NT * 4);  // ring: per slot NA x (double2 sums) then NT x 4 bytes of u16 counts
    b += (size_t)(n_slots + 4) * sizeof(SlotInfo);
    b += (size_t)(NT / 32) * CHUNK * sizeof(short2);
    b += ((size_t)lut_n * 2 + 15) / 16 * 16;
    b += (size_t)(n_slots + 4) * 16;                             // high-word slot table
    b += (size_t)(n_slots + 4) * 4;                              // fp32 screening: lower band edge of the next slot
    return b;
}

#ifndef PIB_V6_CMP
#define PIB_V6_CMP 0
#endif
#ifndef PIB_V6_PIPE
#define PIB_V6_PIPE 0
#endif
#ifndef PIB_V6_HIW
#define PIB_V6_HIW 1
#endif
#ifndef PIB_V6_ANYISSUE
#define PIB_V6_ANYISSUE 0
#endif
#ifndef PIB_F32_LOOKUP2
#define PIB_F32_LOOKUP2 0
#endif
#ifndef PIB_F32_CHAIN
#define PIB_F32_CHAIN 1
#endif
#ifndef PIB_F32_CLAMP
#define PIB_F32_CLAMP 1
#endif
#ifndef PIB_F32_ZRELOAD
#define PIB_F32_ZRELOAD 0
#endif

// packed single precision (two lanes per 64-bit register): FADD2 / FMUL2 / FFMA2 of sm_100
static __device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
static __device__ __forceinline__ unsigned long long f2_pack(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
static __device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <int RING, int NT, int APT, int U, int NL, bool MULTI_FIX, bool ELLIPSE>
__global__ void __launch_bounds__(NT, 1) pair_ring6_kernel(const PairParams p)
{
    constexpr int NA = NT * APT;                   // A points per CTA
    constexpr int A_TILES = NA / TILE;
    constexpr int NW = NT / 32;
    constexpr int WA = 32 * APT;                   // A points per warp (consecutive on the Hilbert curve)
    // 16-bit counts (an A point meets at most CHUNK * TILE = 8192 pairs between folds), the counts of a thread's A points
    // in ONE 32-bit word: a thread stays in its own bank whatever ring block each lane addresses
    constexpr int SLOT_BYTES = NA * 16 + NT * 4;
    constexpr int SB16 = SLOT_BYTES / 16;
    constexpr int CNT_OFF = NA * 16;               // counts follow the sums inside a slot block
    constexpr uint32_t TRASH16 = RING * SB16;
    constexpr int CMP = PIB_V6_CMP;
    constexpr bool PIPE = PIB_V6_PIPE != 0;      // two groups in flight per thread (needs the registers)
    constexpr int NG = TILE / U;
    // high-word fast path (two-lag scheme, plain distances): see sweep_hiw below
    constexpr bool HIW = PIB_V6_HIW != 0 && NL == 2 && !ELLIPSE && U == 8;
    // fp32 screening path (sweep_f32 below): the default two-A-point build only (its tiles take 4 KB of shared memory)
    constexpr bool F32S = HIW && APT == 2 && RING == 10;
    static_assert(NA % TILE == 0 && RING <= 32 && RING >= 4, "A sub-tile is whole tiles; one lane folds one ring slot");
    static_assert(U == 8 || U == 16, "groups of 8 or 16 pairs");
    static_assert(NL == 2 || NL == 4, "a group may span two or four consecutive lags");
    static_assert(NL == 2 || APT == 1, "the four-lag scheme is built for one A point per thread");
    static_assert(APT <= 2, "the counts of a thread's A points share one 32-bit word");
    static_assert((RING + 1) * SB16 < 65536 && SLOT_BYTES % 16 == 0, "ring offsets are 16-bit, in 16-byte units");
    constexpr int STAGES = 4;                      // B tiles in flight: the warps of a CTA may drift apart by STAGES - 1 tiles
    __shared__ __align__(16) double2 s_txy[STAGES][TILE];
    __shared__ __align__(16) double s_tz[STAGES][TILE];
    __shared__ __align__(16) double2 s_gs[STAGES][TILE / 8];
    __shared__ __align__(16) unsigned char s_tf[F32S ? STAGES : 1][F32S ? TF32_TILE_BYTES : 16];
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES];     // TMA landed / every warp is done with the buffer
    __shared__ uint64_t s_live[2];                 // live B tiles of the item (double-buffered over consecutive items)
    __shared__ unsigned s_issued[2];               // B tile loads of the item requested so far (any warp may request the next one)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_slots = p.n_slots, n_lags = n_slots - 2;

    unsigned char *const ring = smem_raw;
    SlotInfo *const s_slot = reinterpret_cast<SlotInfo *>(smem_raw + (size_t)(RING + 1) * SLOT_BYTES);
    short2 *const s_range = reinterpret_cast<short2 *>(s_slot + n_slots + 4) + warp * CHUNK;    // my warp's row
    const uint16_t *const s_lut = reinterpret_cast<const uint16_t *>(reinterpret_cast<short2 *>(s_slot + n_slots + 4) + NW * CHUNK);
    // high-word fast path: {high word of the slot's threshold, high word of the next one, ring byte offsets of this and the next slot}
    int4 *const s_sloth = reinterpret_cast<int4 *>(const_cast<unsigned char *>(reinterpret_cast<const unsigned char *>(s_lut)) + ((size_t)p.lut2_n * 2 + 15) / 16 * 16);
    float *const s_cnext = reinterpret_cast<float *>(s_sloth + n_slots + 4);
    const uint32_t slot_sa = smem_u32(s_slot), lut_sa = smem_u32(s_lut), sloth_sa = smem_u32(s_sloth), cnext_sa = smem_u32(s_cnext);

    // my ring entries: sums at ring + 16 * off + my_acc[k], counts at ring + 16 * off + my_cnt[k]
    int my_acc[APT], my_cnt[APT];
#pragma unroll
    for (int k = 0; k < APT; ++k) {
        my_acc[k] = (k * NT + tid) * 16;
        my_cnt[k] = CNT_OFF + tid * 4 + k * 2;
    }
    uint32_t acc_sa[APT], cnt_sa[APT];             // the same as 32-bit shared addresses of ring block 0
#pragma unroll
    for (int k = 0; k < APT; ++k) {
        acc_sa[k] = smem_u32(ring) + (uint32_t)my_acc[k];
        cnt_sa[k] = smem_u32(ring) + (uint32_t)my_cnt[k];
    }
    for (int s = 0; s <= RING; ++s) {
#pragma unroll
        for (int k = 0; k < APT; ++k) {
            *reinterpret_cast<double2 *>(ring + s * SLOT_BYTES + my_acc[k]) = make_double2(0.0, 0.0);
            *reinterpret_cast<uint16_t *>(ring + s * SLOT_BYTES + my_cnt[k]) = 0;
        }
    }
    for (int s = tid; s < n_slots + 4; s += NT) {          // (entries behind the last slot: +inf thresholds, trash block)
        SlotInfo e;
        e.T = s < n_slots ? __double_as_longlong(p.upper[s]) : 0x7ff0000000000000ll;
        e.next_hi = (int)(__double_as_longlong(p.upper[min(s + 1, n_slots - 1)]) >> 32);
        const uint32_t o_lo = (s >= 1 && s <= n_lags) ? (uint32_t)(s % RING) * SB16 : TRASH16;
        const uint32_t o_hi = (s + 1 >= 1 && s + 1 <= n_lags) ? (uint32_t)((s + 1) % RING) * SB16 : TRASH16;
        e.offs = o_lo | (o_hi << 16);
        s_slot[s] = e;
        if (F32S && p.tf32) {
            const float2 tf = p.slotf[s], tn = p.slotf[min(s + 1, n_slots + 3)];
            s_sloth[s] = make_int4(__float_as_int(tf.x), __float_as_int(tf.y), (int)(o_lo * 16u), (int)(o_hi * 16u));
            // largest float <= (next threshold - its half-width): a squared distance below it is surely inside the next slot's
            // lower band edge
            float c = __fsub_rd(tn.x, tn.y);
            s_cnext[s] = c;
        } else
        s_sloth[s] = make_int4((int)(e.T >> 32), e.next_hi, (int)(o_lo * 16u), (int)(o_hi * 16u));
    }
    const bool f32_on = F32S && p.tf32 != nullptr;
    // (same cells as the fp64 high-word table: a normal float's bits are ((high word - 896 * 2^20) << 3) | 3 more mantissa bits)
    const int base_f = p.lut2_base_f;
    {
        uint16_t *lut_w = const_cast<uint16_t *>(s_lut);
        for (int s = tid; s < p.lut2_n; s += NT) lut_w[s] = p.lut2[s];
    }
    if (tid == 0) {
        for (int b = 0; b < STAGES; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_empty[b], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_tiles = p.n_tiles;
    const int n_atiles = (n_tiles + A_TILES - 1) / A_TILES;
    const int chunk = p.chunk;                     // B tiles per work item (<= CHUNK; smaller for many ranks: finer load balance)
    const int n_chunks = (n_tiles + chunk - 1) / chunk;
    const int64_t n_items = (int64_t)n_atiles * n_chunks;
    uint32_t seq = 0, pseq = 0;                    // B tiles consumed by my warp / B tile loads issued (thread 0), over the whole kernel
    uint32_t item_no = 0;
    unsigned long long evaluated = 0;              // lane 0 of every warp counts its warp's pairs
    double *my_partial = p.partial + ((size_t)blockIdx.x * NW + warp) * n_slots * 4;
    const int base_hi = p.lut2_base, lut_last = p.lut2_n - 1;
    const double w00 = p.w00, w01 = p.w01, w10 = p.w10, w11 = p.w11;
    const double slack = p.box_slack;
    const uint32_t TILE_TX = TILE * (sizeof(double2) + sizeof(double)) + (TILE / U) * sizeof(double2) + (f32_on ? TF32_TILE_BYTES : 0);

    const int64_t n_work = p.item_list ? (int64_t)*p.n_live_items : n_items;
    // The cost-ordered list is dealt to the (rank, CTA) slots in rounds, in SNAKE order: dealt in the same direction every round,
    // slot 0 gets the most expensive item of each round and ends up one whole item above the last slot (3.6 % of a rank's
    // kernel time at 8 ranks); alternating the direction cancels that round by round.  Static, so a warp's partial sums are
    // the same in every run.
    const int64_t deal = (int64_t)gridDim.x * p.shard_count, my_slot = (int64_t)blockIdx.x * p.shard_count + p.shard_index;
    for (int64_t round = 0; round * deal < n_work; ++round) {
        const int64_t kk = round * deal + ((round & 1) ? deal - 1 - my_slot : my_slot);
        if (kk >= n_work) continue;
        const int64_t item = p.item_list ? (int64_t)p.item_list[kk] : kk;
        const int64_t a0 = (item / n_chunks) * NA;            // first point of the CTA's A sub-tile
        const int ta = (int)(a0 / TILE);
        const int ta_last = min(ta + A_TILES, n_tiles) - 1;
        const int tb0 = (int)(item % n_chunks) * chunk;
        const int tb1 = min(tb0 + chunk, n_tiles);
        if (tb1 <= ta || a0 >= p.n) continue;
        const int tb_first = max(tb0, ta);

        // ---- my A points -----------------------------------------------------------------------
        int ig[APT];
        double xi[APT], yi[APT], zi[APT];
        double bx0 = INFINITY, by0 = INFINITY, bx1 = -INFINITY, by1 = -INFINITY;     // my warp's box
#pragma unroll
        for (int k = 0; k < APT; ++k) {
            ig[k] = (int)a0 + warp * WA + k * 32 + lane;
            // the last A sub-tile may reach past the padded arrays (n_pad is a multiple of TILE, not of NA): rows >= n never
            // count (every B tile they meet is a diagonal tile and keeps j > i only), but their loads must stay inside
            const int il = min(ig[k], n_tiles * TILE - 1);
            PIB_DBG_ASSERT(il >= 0 && il < n_tiles * TILE);
            const double2 pa = p.xy[il];
            xi[k] = pa.x; yi[k] = pa.y; zi[k] = p.z[il];
            if (ig[k] < p.n) {
                const double u = ELLIPSE ? w00 * pa.x + w01 * pa.y : pa.x, v = ELLIPSE ? w10 * pa.x + w11 * pa.y : pa.y;
                bx0 = fmin(bx0, u); bx1 = fmax(bx1, u); by0 = fmin(by0, v); by1 = fmax(by1, v);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, o));
        }

        // ---- CTA: which B tiles of the chunk are loaded at all (box of the whole A sub-tile) -----------
        if (warp == 0) {
            double4 abox = p.tile_box[ta];
            for (int t = ta + 1; t <= ta_last; ++t) {
                const double4 o = p.tile_box[t];
                abox.x = fmin(abox.x, o.x); abox.y = fmin(abox.y, o.y);
                abox.z = fmax(abox.z, o.z); abox.w = fmax(abox.w, o.w);
            }
            uint64_t m = 0;
            for (int h = 0; h < (chunk + 31) / 32; ++h) {
                const int tb = tb0 + h * 32 + lane;
                bool live = false;
                if (tb >= tb_first && tb < tb1) {
                    const double4 bbox = p.tile_box[tb];
                    double dmin;
                    if (ELLIPSE) {
                        const double gx = fmax(0.0, fmax(abox.x - bbox.z, bbox.x - abox.z) - slack);
                        const double gy = fmax(0.0, fmax(abox.y - bbox.w, bbox.y - abox.w) - slack);
                        dmin = (gx * gx + gy * gy) * (1.0 - 1e-12);
                    } else dmin = box_min_d2(abox, bbox);
                    live = dmin <= p.max_d2;
                }
                m |= (uint64_t)__ballot_sync(0xffffffffu, live) << (32 * h);
            }
            if (lane == 0) { s_live[item_no & 1] = m; s_issued[item_no & 1] = 0; }
        }
        // ---- my warp: the lag slots each B tile can touch for MY A points (empty = skip the tile) ----------
        uint64_t f32_tiles = 0;                               // B tiles of the chunk the fp32 screening path may take
        for (int h = 0; h < (chunk + 31) / 32; ++h) {
            const int tb = tb0 + h * 32 + lane;
            short2 r = make_short2(1, 0);                     // empty
            bool narrow = false;
            if (tb >= tb_first && tb < tb1) {
                const double4 bbox = p.tile_box[tb];
                // full tiles whose points lie within f32_rcap of the tile's origin (the error bound of the screening)
                narrow = f32_on && (int64_t)(tb + 1) * TILE <= p.n &&
                         0.5 * fmax(bbox.z - bbox.x, bbox.w - bbox.y) <= p.f32_rcap;
                const double sl = ELLIPSE ? slack : 0.0;
                const double gx = fmax(0.0, fmax(bx0 - bbox.z, bbox.x - bx1) - sl);
                const double gy = fmax(0.0, fmax(by0 - bbox.w, bbox.y - by1) - sl);
                // every pair of the two boxes has dmin <= d2 <= dmax in the kernel's own arithmetic (fp subtraction,
                // multiplication and addition are monotone; the whitened bounds carry slack and a relative margin)
                const double dmin = ELLIPSE ? (gx * gx + gy * gy) * (1.0 - 1e-12) : __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
                if (dmin <= p.max_d2) {
                    const double dxm = fmax(bx1 - bbox.x, bbox.z - bx0) + sl, dym = fmax(by1 - bbox.y, bbox.w - by0) + sl;
                    const double dmax = ELLIPSE ? (dxm * dxm + dym * dym) * (1.0 + 1e-12)
                                                : __dadd_rn(__dmul_rn(dxm, dxm), __dmul_rn(dym, dym));
                    int a = 0, b = n_slots - 1;               // first slot whose threshold is >= dmin
                    const long long v = __double_as_longlong(dmin);
                    while (a < b) { const int mid = (a + b) >> 1; if (v <= s_slot[mid].T) b = mid; else a = mid + 1; }
                    const int lo = max(a, 1);
                    a = 0; b = n_slots - 1;
                    const long long w = __double_as_longlong(dmax);
                    while (a < b) { const int mid = (a + b) >> 1; if (w <= s_slot[mid].T) b = mid; else a = mid + 1; }
                    r = make_short2((short)lo, (short)min(a, n_lags));
                }
            }
            s_range[h * 32 + lane] = r;
            if (F32S) f32_tiles |= (uint64_t)__ballot_sync(0xffffffffu, narrow) << (32 * h);
        }
        __syncthreads();                                      // the one CTA-wide barrier of an item (s_live is double-buffered)
        uint64_t live = s_live[item_no & 1];
        const uint32_t item_slot = item_no & 1;
        ++item_no;
        if (live == 0) continue;

        // My warp's ring window: lag slots [w_lo, w_lo + RING), slot s in ring block s % RING; w_lo < 0 = every block is empty.
        // The window SLIDES: when a tile needs slots beyond it, only the blocks that leave the window are folded into the
        // warp's global partials (about half a block per tile on BASELINE config 2 with RING = 10, against 1.7 for
        // folding the whole ring each time the union of the tiles' ranges outgrew it).
        int w_lo = -1;

        // fold the ring blocks of slots [s_from, s_to) (inside the window).  Every thread turns ITS A points' entries of a block
        // into the block's four sums (sum dz^2 = c z^2 - 2 z S1 + S2;  sum z_i z_j = z S1;  sum (z_i + z_j) = c z + S1), a
        // butterfly adds them over the warp in a fixed order, and lane `block` adds the totals to the warp's partials.  (The
        // first version let lane r walk block r over the warp's points: a dependent shared-memory chain per point that cost
        // about as much as a whole B tile.)
        auto evict = [&](int s_from, int s_to) {
            __syncwarp();
#pragma unroll 1
            for (int sl = s_from; sl < s_to; ++sl) {
                const int r = sl % RING;
                PIB_DBG_ASSERT(sl >= 0 && sl < n_slots + RING && r >= 0);
                double v_sq = 0.0, v_prod = 0.0, v_val = 0.0;
                unsigned v_cnt = 0;
#pragma unroll
                for (int k = 0; k < APT; ++k) {
                    double2 *ap = reinterpret_cast<double2 *>(ring + r * SLOT_BYTES + my_acc[k]);
                    uint16_t *cp = reinterpret_cast<uint16_t *>(ring + r * SLOT_BYTES + my_cnt[k]);
                    const double2 a = *ap;
                    const unsigned c = *cp;
                    *ap = make_double2(0.0, 0.0);
                    *cp = 0;
                    const double cd = (double)c, z = zi[k];
                    v_sq += __fma_rn(z, __fma_rn(cd, z, -2.0 * a.x), a.y);
                    v_prod = __fma_rn(z, a.x, v_prod);
                    v_val += __fma_rn(cd, z, a.x);
                    v_cnt += c;
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    v_sq += __shfl_xor_sync(0xffffffffu, v_sq, o);
                    v_prod += __shfl_xor_sync(0xffffffffu, v_prod, o);
                    v_val += __shfl_xor_sync(0xffffffffu, v_val, o);
                    v_cnt += __shfl_xor_sync(0xffffffffu, v_cnt, o);
                }
                if (lane == r && v_cnt) {
                    // fire-and-forget adds: a slot always lives in the same ring block, so the same lane issues all adds to an
                    // address, in program order — the sums stay reproducible run to run
                    PIB_DBG_ASSERT(sl >= 1 && sl <= n_lags);
                    double *dst = my_partial + (size_t)sl * 4;
                    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst), "d"(v_sq) : "memory");
                    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 1), "d"(v_prod) : "memory");
                    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 2), "d"(v_val) : "memory");
                    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst + 3), "l"((unsigned long long)v_cnt) : "memory");
                }
            }
            __syncwarp();
        };
        auto fold = [&]() {
            if (w_lo >= 0) evict(w_lo, w_lo + RING);
            w_lo = -1;
        };

        // B tiles stream through STAGES buffers: thread 0 issues the loads STAGES - 1 tiles ahead (waiting for every warp to
        // have released the buffer's previous tile), each warp waits for a tile to land, sweeps it and releases the buffer.
        // No CTA-wide barrier per tile: warps that cull a tile, or finish it early, run ahead.
        auto issue_tile = [&](int t) {
            PIB_DBG_ASSERT(t >= 0 && t < n_tiles);
            const uint32_t b = pseq % STAGES;
            if (pseq >= STAGES) mbar_wait(&s_empty[b], ((pseq / STAGES) - 1) & 1);
            mbar_expect_tx(&s_full[b], TILE_TX);
            tma_load_1d(s_txy[b], p.xy + (size_t)t * TILE, TILE * sizeof(double2), &s_full[b]);
            tma_load_1d(s_tz[b], p.z + (size_t)t * TILE, TILE * sizeof(double), &s_full[b]);
            tma_load_1d(s_gs[b], p.gsum + (size_t)t * (TILE / U), (TILE / U) * sizeof(double2), &s_full[b]);
            if (F32S && f32_on) tma_load_1d(s_tf[b], p.tf32 + (size_t)t * TF32_TILE_BYTES, TF32_TILE_BYTES, &s_full[b]);
            ++pseq;
        };
#if PIB_V6_ANYISSUE
        // Any warp requests the next loads: a buffer is refilled as soon as its last reader has released it (by that reader, on
        // its way into the next tile) instead of when one fixed producer thread comes by — with thread 0 as the only producer
        // warp 0 waited at every tile for the slowest warp to release the previous one, and everybody else for warp 0.
        const uint64_t live_all = live;
        const int n_live = __popcll(live_all);
        const uint32_t seq0 = seq;                            // (every warp consumes every live tile: CTA-uniform at the item's start)
        auto request_tiles = [&]() {                          // lane 0 of any warp
            volatile unsigned *issued = &s_issued[item_slot];
            while (true) {
                const unsigned t = *issued;
                if ((int)t >= n_live) break;
                const uint32_t g = seq0 + t, b = g % STAGES;
                if (g >= STAGES && !mbar_test(&s_empty[b], ((g / STAGES) - 1) & 1)) break;     // a reader still holds the buffer
                if (atomicCAS(&s_issued[item_slot], t, t + 1) != t) continue;                  // another warp took this one
                // the t-th live tile of the item
                const uint32_t m_lo = (uint32_t)live_all, m_hi = (uint32_t)(live_all >> 32);
                const unsigned c_lo = __popc(m_lo);
                const int bit = t < c_lo ? (int)__fns(m_lo, 0, (int)t + 1) : 32 + (int)__fns(m_hi, 0, (int)(t - c_lo) + 1);
                const int tile = tb0 + bit;
                PIB_DBG_ASSERT(bit >= 0 && bit < 64 && tile >= 0 && tile < n_tiles);
                mbar_expect_tx(&s_full[b], TILE_TX);
                tma_load_1d(s_txy[b], p.xy + (size_t)tile * TILE, TILE * sizeof(double2), &s_full[b]);
                tma_load_1d(s_tz[b], p.z + (size_t)tile * TILE, TILE * sizeof(double), &s_full[b]);
                tma_load_1d(s_gs[b], p.gsum + (size_t)tile * (TILE / U), (TILE / U) * sizeof(double2), &s_full[b]);
                if (F32S && f32_on) tma_load_1d(s_tf[b], p.tf32 + (size_t)tile * TF32_TILE_BYTES, TF32_TILE_BYTES, &s_full[b]);
            }
        };
        (void)pseq; (void)issue_tile;
        while (live) {
            const int tb = tb0 + __ffsll((long long)live) - 1;
            live &= live - 1;
            if (lane == 0) request_tiles();
            __syncwarp();
#else
        uint64_t ahead = live;                                // thread 0: tiles not yet requested
        if (tid == 0) {
            for (int i = 0; i < STAGES - 1 && ahead; ++i) {
                issue_tile(tb0 + __ffsll((long long)ahead) - 1);
                ahead &= ahead - 1;
            }
        }
        while (live) {
            const int tb = tb0 + __ffsll((long long)live) - 1;
            live &= live - 1;
            if (tid == 0 && ahead) {
                issue_tile(tb0 + __ffsll((long long)ahead) - 1);
                ahead &= ahead - 1;
            }
#endif
            const uint32_t b = seq % STAGES, phase = (seq / STAGES) & 1;
            // ---- my warp's ring window (warp-uniform) ---------------------------------------------------
            const short2 rng = s_range[tb - tb0];
            const bool mine = rng.y >= rng.x;                 // can any of my warp's points reach this tile?
            const bool wide = mine && (rng.y - rng.x >= RING);
            if (mine && !wide) {
                const int l = rng.x, h = rng.y;
                if (w_lo < 0) w_lo = max(1, l - (RING - (h - l + 1)) / 2);       // empty ring: the tile's range in the middle
                else if (l < w_lo) { evict(max(l + RING, w_lo), w_lo + RING); w_lo = l; }
                else if (h >= w_lo + RING) { evict(w_lo, min(h - RING + 1, w_lo + RING)); w_lo = h - RING + 1; }
            }
            mbar_wait(&s_full[b], phase);
            ++seq;

            if (mine) {
                const double2 *bxy = s_txy[b];
                const double *bz = s_tz[b];
                const double2 *bgs = s_gs[b];
                const bool diagonal = (tb <= ta_last);
                const int jbase = tb * TILE;
                {
                    // pairs (i, j) with j > i of my warp against this tile
                    const int jend = (int)min((int64_t)jbase + TILE, p.n);
                    unsigned np = 0;
#pragma unroll
                    for (int k = 0; k < APT; ++k)
                        if (ig[k] < p.n) np += (unsigned)max(0, jend - max(ig[k] + 1, jbase));
                    np = __reduce_add_sync(0xffffffffu, np);
                    evaluated += np;
                }
                // fp32 screening: my A points relative to the tile's origin, each duplicated into both lanes of a pair
                const bool tile_f32 = F32S && !diagonal && ((f32_tiles >> (tb - tb0)) & 1);
                unsigned long long xa2[APT], ya2[APT];
#pragma unroll
                for (int k = 0; k < APT; ++k) xa2[k] = ya2[k] = 0;
                if (F32S && tile_f32) {
                    const double2 org = *reinterpret_cast<const double2 *>(s_tf[F32S ? b : 0] + TILE * 8);
#pragma unroll
                    for (int k = 0; k < APT; ++k) {
                        const float xf = __double2float_rn(__dsub_rn(xi[k], org.x)), yf = __double2float_rn(__dsub_rn(yi[k], org.y));
                        xa2[k] = f2_pack(xf, xf);
                        ya2[k] = f2_pack(yf, yf);
                    }
                }
                int wlo = 1;                                  // wide sweeps: lag slots the ring accepts, [wlo, wlo + wlen)
                unsigned wlen = (unsigned)n_lags;

                // squared distance in the reference's operation order
                auto dist2 = [&](const double2 pj, int k) __attribute__((always_inline)) -> double {
                    if (ELLIPSE) {
                        // vector_distance = other - centre (angular.py:448); W @ d as OpenBLAS dgemm evaluates it:
                        // fma(W[r][1], dy, W[r][0]*dx); einsum: n0*n0 + n1*n1 un-fused (angular.py:666-668)
                        const double dx = __dsub_rn(pj.x, xi[k]), dy = __dsub_rn(pj.y, yi[k]);
                        const double n0 = __fma_rn(w01, dy, __dmul_rn(w00, dx));
                        const double n1 = __fma_rn(w11, dy, __dmul_rn(w10, dx));
                        return __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
                    } else {
                        // scipy cdist euclidean: sqrt(dx*dx + dy*dy), nothing fused (distance/point.py:56)
                        const double dx = __dsub_rn(xi[k], pj.x), dy = __dsub_rn(yi[k], pj.y);
                        return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                };
                // one pair into the ring (slow paths)
                auto add_pair = [&](int q, int k, double z, bool wide_mode) __attribute__((always_inline)) {
                    PIB_DBG_ASSERT(q >= 0 && q < n_slots + 4);
                    uint32_t off = s_slot[q].offs & 0xffffu;
                    if (wide_mode && !((unsigned)(q - wlo) < wlen)) off = TRASH16;
                    double2 *ap = reinterpret_cast<double2 *>(ring + off * 16 + my_acc[k]);
                    uint16_t *cp = reinterpret_cast<uint16_t *>(ring + off * 16 + my_cnt[k]);
                    double2 a = *ap;
                    a.x += z; a.y = __fma_rn(z, z, a.y);
                    *ap = a;
                    *cp = (uint16_t)(*cp + 1u);
                };

                // ---- one group of U B points, in two stages so that two groups are in flight ----------------------
                struct Grp {
                    long long d2[APT][U];
                    long long T[APT];          // threshold between the group's first and second lag
                    uint32_t offs[APT];        // ring blocks of the two lags
                    uint32_t ea[APT];          // shared address of the first lag's table entry
                    bool beyond;               // a pair may lie beyond the group's last lag
                    long long T1, T2;          // NL == 4: the thresholds behind the second and third lag
                    uint32_t offs23;           //          ring blocks of the third and fourth lag
                };
                // stage A: the U squared distances of every A point, the group's first lag slot and its table entry
                auto stage_a = [&](int g, Grp &s) __attribute__((always_inline)) {
                    const int j0 = g * U;
                    int mn[APT], mx[APT];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const double2 pj = bxy[j0 + u];
#pragma unroll
                        for (int k = 0; k < APT; ++k) {
                            s.d2[k][u] = __double_as_longlong(dist2(pj, k));
                            const int h = (int)(s.d2[k][u] >> 32);
                            mn[k] = u ? min(mn[k], h) : h;
                            mx[k] = u ? max(mx[k], h) : h;
                        }
                    }
                    // the smallest high word with a zero low word is a lower bound of all U squared distances; its slot is
                    // <= every pair's slot (== the smallest unless a threshold falls into the 2^-20 relative gap, which
                    // only sends more pairs to the slow path)
                    s.beyond = false;
#pragma unroll
                    for (int k = 0; k < APT; ++k) {
                        int idx = (mn[k] - base_hi) >> LUT2_SHIFT;
                        idx = min(max(idx, 0), lut_last);
                        unsigned short q0;
                        asm("ld.shared.u16 %0, [%1];" : "=h"(q0) : "r"(lut_sa + 2u * (uint32_t)idx));
                        PIB_DBG_ASSERT((int)q0 < n_slots);
                        uint32_t ea = slot_sa + 16u * (uint32_t)q0;
                        int thi;
                        if (MULTI_FIX) {
                            while (true) {
                                asm volatile("ld.shared.s32 %0, [%1+4];" : "=r"(thi) : "r"(ea));
                                if (!(mn[k] > thi)) break;
                                ea += 16u;
                            }
                        } else {
                            asm("ld.shared.s32 %0, [%1+4];" : "=r"(thi) : "r"(ea));
                            ea += (mn[k] > thi) ? 16u : 0u;
                        }
                        int e_tlo, e_thi, e_next;
                        asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(e_tlo), "=r"(e_thi), "=r"(e_next), "=r"(s.offs[k]) : "r"(ea));
                        s.T[k] = ((long long)e_thi << 32) | (unsigned)e_tlo;
                        s.ea[k] = ea;
                        if (NL == 4) {
                            // four lags per group: thresholds of the next two slots, the ring blocks of slots q + 2, q + 3 and
                            // the high word of the threshold behind the fourth lag
                            int t1lo, t1hi, t2lo, t2hi, n2;
                            asm("ld.shared.v2.s32 {%0, %1}, [%2+16];" : "=r"(t1lo), "=r"(t1hi) : "r"(ea));
                            asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4+32];" : "=r"(t2lo), "=r"(t2hi), "=r"(n2), "=r"(s.offs23) : "r"(ea));
                            s.T1 = ((long long)t1hi << 32) | (unsigned)t1lo;
                            s.T2 = ((long long)t2hi << 32) | (unsigned)t2lo;
                            e_next = n2;
                        }
                        s.beyond |= (mx[k] >= e_next);
                    }
                };
                // second-lag sums of one group for A point k and the two ring updates; g1, g2, n_in = sums of z, z^2 and
                // the number of pairs the ring still has to take (the slow path removes pairs it added on its own)
                auto finish_group = [&](int k, int j0, const long long (&d2k)[U], long long T, uint32_t offs, int q, double g1,
                                        double g2, int n_in, const bool wide_mode, const bool bounded, long long t1)
                                        __attribute__((always_inline)) {
                    uint32_t o_lo = offs & 0xffffu, o_hi = offs >> 16;
                    PIB_DBG_ASSERT(o_lo <= TRASH16 && o_hi <= TRASH16 && o_lo % SB16 == 0 && o_hi % SB16 == 0);
                    if (wide_mode) {
                        if (!((unsigned)(q - wlo) < wlen)) o_lo = TRASH16;
                        if (!((unsigned)(q + 1 - wlo) < wlen)) o_hi = TRASH16;
                    }
                    double h1a, h1b, h2a, h2b;                // (started from the first pair: x + 0.0 is not free in fp64)
                    int ch = 0;
#pragma unroll
                    for (int u = 0; u < U; u += 2) {
                        const double2 zz = *reinterpret_cast<const double2 *>(bz + j0 + u);
                        // "second lag" predicate on the bit patterns (d2 >= 0: integer order == fp64 order); one select per
                        // pair: a first-lag pair keeps the low word of z under a zero high word, a subnormal below 1e-308
                        // whose contribution to the fp64 sums is far below one ulp; predicated count
                        int za_hi, zb_hi;
                        if (bounded) {
                            // slow variant: pairs beyond the second lag (d2 > t1) were added on their own and stay out
                            const bool pa = d2k[u] > T && !(d2k[u] > t1), pb = d2k[u + 1] > T && !(d2k[u + 1] > t1);
                            za_hi = pa ? __double2hiint(zz.x) : 0; zb_hi = pb ? __double2hiint(zz.y) : 0;
                            ch += (pa ? 1 : 0) + (pb ? 1 : 0);
                        } else {
                        if (CMP == 1)
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(za_hi), "+r"(ch) : "r"(__double2hiint(zz.x)), "d"(__longlong_as_double(d2k[u])), "d"(__longlong_as_double(T)));
                        else
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.s64 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(za_hi), "+r"(ch) : "r"(__double2hiint(zz.x)), "l"(d2k[u]), "l"(T));
                        if (CMP >= 1)
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(zb_hi), "+r"(ch) : "r"(__double2hiint(zz.y)), "d"(__longlong_as_double(d2k[u + 1])), "d"(__longlong_as_double(T)));
                        else
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.s64 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(zb_hi), "+r"(ch) : "r"(__double2hiint(zz.y)), "l"(d2k[u + 1]), "l"(T));
                        }
                        const double za = __hiloint2double(za_hi, __double2loint(zz.x));
                        const double zb = __hiloint2double(zb_hi, __double2loint(zz.y));
                        if (u == 0) {
                            h1a = za; h2a = __dmul_rn(za, za);
                            h1b = zb; h2b = __dmul_rn(zb, zb);
                        } else {
                            h1a += za; h2a = __fma_rn(za, za, h2a);
                            h1b += zb; h2b = __fma_rn(zb, zb, h2b);
                        }
                    }
                    const double h1 = h1a + h1b, h2 = h2a + h2b;
                    double2 *alo = reinterpret_cast<double2 *>(ring + o_lo * 16 + my_acc[k]);
                    double2 *ahi = reinterpret_cast<double2 *>(ring + o_hi * 16 + my_acc[k]);
                    uint16_t *clo = reinterpret_cast<uint16_t *>(ring + o_lo * 16 + my_cnt[k]);
                    uint16_t *chi = reinterpret_cast<uint16_t *>(ring + o_hi * 16 + my_cnt[k]);
                    double2 v_lo = *alo, v_hi = *ahi;
                    const uint32_t c_lo = *clo, c_hi = *chi;
                    v_lo.x += g1 - h1; v_lo.y += g2 - h2;
                    v_hi.x += h1;      v_hi.y += h2;
                    *alo = v_lo; *ahi = v_hi;               // both blocks are the trash block when they coincide: never read
                    *clo = (uint16_t)(c_lo + (uint32_t)(n_in - ch));
                    *chi = (uint16_t)(c_hi + (uint32_t)ch);
                };
                // NL == 4: cumulative sums over the pairs above each of the three thresholds; slot q + i gets the difference of
                // two consecutive ones (q: all - H0, q + 1: H0 - H1, q + 2: H1 - H2, q + 3: H2)
                auto finish_group4 = [&](int j0, const long long (&d2k)[U], const Grp &s, int q, double g1, double g2, int n_in,
                                         const bool wide_mode, const bool bounded, long long t_last) __attribute__((always_inline)) {
                    uint32_t o[4] = {s.offs[0] & 0xffffu, s.offs[0] >> 16, s.offs23 & 0xffffu, s.offs23 >> 16};
                    if (wide_mode) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (!((unsigned)(q + i - wlo) < wlen)) o[i] = TRASH16;
                    }
                    const long long T[3] = {s.T[0], s.T1, s.T2};
                    double h1[3] = {0.0, 0.0, 0.0}, h2[3] = {0.0, 0.0, 0.0};
                    int c[3] = {0, 0, 0};
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const double zj = bz[j0 + u];
                        const bool in = !bounded || !(d2k[u] > t_last);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            int zh;
                            if (bounded) {
                                const bool pr = in && d2k[u] > T[i];
                                zh = pr ? __double2hiint(zj) : 0;
                                c[i] += pr ? 1 : 0;
                            } else {
                                asm("{\n\t.reg .pred p;\n\tsetp.gt.s64 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                    : "=r"(zh), "+r"(c[i]) : "r"(__double2hiint(zj)), "l"(d2k[u]), "l"(T[i]));
                            }
                            const double zm = __hiloint2double(zh, __double2loint(zj));
                            h1[i] += zm; h2[i] = __fma_rn(zm, zm, h2[i]);
                        }
                    }
                    const double a1[4] = {g1 - h1[0], h1[0] - h1[1], h1[1] - h1[2], h1[2]};
                    const double a2[4] = {g2 - h2[0], h2[0] - h2[1], h2[1] - h2[2], h2[2]};
                    const int n[4] = {n_in - c[0], c[0] - c[1], c[1] - c[2], c[2]};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // one slot after the other: two of the four may be the trash block
                        double2 *ap = reinterpret_cast<double2 *>(ring + o[i] * 16 + my_acc[0]);
                        uint16_t *cp = reinterpret_cast<uint16_t *>(ring + o[i] * 16 + my_cnt[0]);
                        double2 v = *ap;
                        v.x += a1[i]; v.y += a2[i];
                        *ap = v;
                        *cp = (uint16_t)(*cp + (uint32_t)n[i]);
                    }
                };
                // stage B: sums and ring updates
                auto stage_b = [&](int g, const Grp &s, const bool wide_mode) __attribute__((always_inline)) {
                    const int j0 = g * U;
                    const double2 gs = bgs[g];
                    if (!__any_sync(0xffffffffu, s.beyond)) {
                        // fast path: every pair of the warp is in one of the group's NL lags
                        if (NL == 4) {
                            finish_group4(j0, s.d2[0], s, wide_mode ? (int)((s.ea[0] - slot_sa) >> 4) : 0, gs.x, gs.y, U, wide_mode, false, 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < APT; ++k)
                                finish_group(k, j0, s.d2[k], s.T[k], s.offs[k], wide_mode ? (int)((s.ea[k] - slot_sa) >> 4) : 0, gs.x, gs.y,
                                             U, wide_mode, false, 0);
                        }
                    } else {
                        // rare: a pair may lie beyond the group's last lag — those pairs are added on their own and leave the
                        // group's sums (the distances stay read-only: both branches use the same registers)
#pragma unroll
                        for (int k = 0; k < APT; ++k) {
                            const int q = (int)((s.ea[k] - slot_sa) >> 4);
                            const long long t1 = s_slot[q + NL - 1].T;
                            double g1 = gs.x, g2 = gs.y;
                            int n_in = U;
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                if (s.d2[k][u] > t1) {
                                    int sx = q + NL;
                                    while (s.d2[k][u] > s_slot[sx].T) ++sx;
                                    const double zj = bz[j0 + u];
                                    add_pair(sx, k, zj, wide_mode);
                                    g1 -= zj; g2 = __fma_rn(-zj, zj, g2);
                                    --n_in;
                                }
                            }
                            if (NL == 4) finish_group4(j0, s.d2[k], s, q, g1, g2, n_in, wide_mode, true, t1);
                            else finish_group(k, j0, s.d2[k], s.T[k], s.offs[k], q, g1, g2, n_in, wide_mode, true, t1);
                        }
                    }
                };

                // ---- high-word fast path ------------------------------------------------------------------------------------
                // The squared distance is formed with ONE fused step, d2' = fma(dx, dx, dy * dy) (4 fp64 instructions instead of
                // the 5 of the un-fused reference order), and only its HIGH word is kept: the lag of a pair is decided by
                // hi(d2') > hi(T).  d2' is at most one ulp away from the reference's d2, so whenever hi(d2') != hi(T) both sit
                // on the same side of T and the decision is exactly the reference's (the host clears p.hiw_ok for a threshold
                // whose low word is all ones, the one pattern where one ulp could cross T between two high words).  A group
                // with a pair ON the threshold's high word (relative band 2^-20: ~1 % of the warp-wide groups) is redone by the
                // exact stages above, like a group with a pair beyond its second lag.
                // The stages of consecutive groups are interleaved in ONE basic block — distances + table lookup of group
                // g + 1 (fp64 pipe) beside the compares, selects and ring updates of group g (integer pipe, shared memory) —
                // which the 16 registers freed by dropping the low words pay for.
                struct GrpH {
                    int x[APT][U];             // high words of the squared distances
                    int t_hi[APT];             // high word of the threshold between the group's first and second lag
                    uint32_t off_lo[APT], off_hi[APT];   // ring byte offsets of the group's two lags
                    bool redo;
                };
                auto stage_a_h = [&](int g, GrpH &s) __attribute__((always_inline)) {
                    const int j0 = g * U;
                    int mn[APT], mx[APT];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const double2 pj = bxy[j0 + u];
#pragma unroll
                        for (int k = 0; k < APT; ++k) {
                            const double dx = __dsub_rn(xi[k], pj.x), dy = __dsub_rn(yi[k], pj.y);
                            const int h = __double2hiint(__fma_rn(dx, dx, __dmul_rn(dy, dy)));
                            s.x[k][u] = h;
                            mn[k] = u ? min(mn[k], h) : h;
                            mx[k] = u ? max(mx[k], h) : h;
                        }
                    }
                    s.redo = false;
#pragma unroll
                    for (int k = 0; k < APT; ++k) {
                        int idx = (mn[k] - base_hi) >> LUT2_SHIFT;
                        idx = min(max(idx, 0), lut_last);
                        unsigned short q0;
                        asm("ld.shared.u16 %0, [%1];" : "=h"(q0) : "r"(lut_sa + 2u * (uint32_t)idx));
                        PIB_DBG_ASSERT((int)q0 < n_slots);
                        uint32_t ea = sloth_sa + 16u * (uint32_t)q0;
                        int thi;
                        if (MULTI_FIX) {
                            while (true) {
                                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(thi) : "r"(ea));
                                if (!(mn[k] > thi)) break;
                                ea += 16u;
                            }
                        } else {
                            asm("ld.shared.s32 %0, [%1];" : "=r"(thi) : "r"(ea));
                            ea += (mn[k] > thi) ? 16u : 0u;
                        }
                        int e_next;
                        asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(s.t_hi[k]), "=r"(e_next), "=r"(s.off_lo[k]), "=r"(s.off_hi[k]) : "r"(ea));
                        bool on_edge = false;
#pragma unroll
                        for (int u = 0; u < U; ++u) on_edge |= (s.x[k][u] == s.t_hi[k]);
                        s.redo |= on_edge | (mx[k] >= e_next);
                    }
                };
                auto stage_b_h = [&](int g, const GrpH &s) __attribute__((always_inline)) {
                    const int j0 = g * U;
                    const double2 gs = bgs[g];
#pragma unroll
                    for (int k = 0; k < APT; ++k) {
                        double h1a, h1b, h2a, h2b;
                        int ch = 0;
#pragma unroll
                        for (int u = 0; u < U; u += 2) {
                            const double2 zz = *reinterpret_cast<const double2 *>(bz + j0 + u);
                            int za_hi, zb_hi;
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(za_hi), "+r"(ch) : "r"(__double2hiint(zz.x)), "r"(s.x[k][u]), "r"(s.t_hi[k]));
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(zb_hi), "+r"(ch) : "r"(__double2hiint(zz.y)), "r"(s.x[k][u + 1]), "r"(s.t_hi[k]));
                            const double za = __hiloint2double(za_hi, __double2loint(zz.x));
                            const double zb = __hiloint2double(zb_hi, __double2loint(zz.y));
                            if (u == 0) {
                                h1a = za; h2a = __dmul_rn(za, za);
                                h1b = zb; h2b = __dmul_rn(zb, zb);
                            } else {
                                h1a += za; h2a = __fma_rn(za, za, h2a);
                                h1b += zb; h2b = __fma_rn(zb, zb, h2b);
                            }
                        }
                        const double h1 = h1a + h1b, h2 = h2a + h2b;
                        // ring updates on 32-bit shared addresses: one add per address (the table holds byte offsets)
                        const uint32_t a_lo = acc_sa[k] + s.off_lo[k], a_hi = acc_sa[k] + s.off_hi[k];
                        const uint32_t n_lo = cnt_sa[k] + s.off_lo[k], n_hi = cnt_sa[k] + s.off_hi[k];
                        PIB_DBG_ASSERT(s.off_lo[k] <= (uint32_t)RING * SLOT_BYTES && s.off_hi[k] <= (uint32_t)RING * SLOT_BYTES &&
                                       s.off_lo[k] % SLOT_BYTES == 0 && s.off_hi[k] % SLOT_BYTES == 0);
                        double2 v_lo, v_hi;
                        uint32_t c_lo, c_hi;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v_lo.x), "=d"(v_lo.y) : "r"(a_lo) : "memory");
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v_hi.x), "=d"(v_hi.y) : "r"(a_hi) : "memory");
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(c_lo) : "r"(n_lo) : "memory");
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(c_hi) : "r"(n_hi) : "memory");
                        v_lo.x += gs.x - h1; v_lo.y += gs.y - h2;
                        v_hi.x += h1;        v_hi.y += h2;
                        c_lo = c_lo + (uint32_t)U - (uint32_t)ch;
                        c_hi = c_hi + (uint32_t)ch;
                        // (both blocks are the trash block when they coincide: never read)
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a_lo), "d"(v_lo.x), "d"(v_lo.y) : "memory");
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a_hi), "d"(v_hi.x), "d"(v_hi.y) : "memory");
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(n_lo), "r"(c_lo) : "memory");
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(n_hi), "r"(c_hi) : "memory");
                    }
                };
                // group g's second stage beside group g + 1's first, in one basic block; a group that must be redone exactly only
                // leaves its bit in `redo` — the exact stages run after the sweep, outside the hot loop's code
                unsigned redo = 0;
                auto step_h = [&](int g, const GrpH &cur, GrpH &nxt) __attribute__((always_inline)) {
                    if (!__any_sync(0xffffffffu, cur.redo)) {
                        // (first stage first: its shared-memory loads must not queue behind the ring stores of the second stage)
                        stage_a_h(g + 1, nxt);
                        stage_b_h(g, cur);
                    } else {
                        redo |= 1u << g;
                        stage_a_h(g + 1, nxt);
                    }
                };
                auto sweep_hiw = [&]() __attribute__((always_inline)) {
                    if (PIB_V6_HIW == 2 || (PIB_V6_HIW == 1 && APT == 2)) {
                        // one group after the other: the smallest loop body.  With two A points per thread the two chains already
                        // overlap and the interleaved form below only costs instruction-cache misses (its body plus prologue and
                        // epilogue is 19 KB against a 6 KB L0: 157.7 ms; this form: 147.7 ms)
#pragma unroll 1
                        for (int g = 0; g < NG; ++g) {
                            GrpH s0;
                            stage_a_h(g, s0);
                            if (!__any_sync(0xffffffffu, s0.redo)) stage_b_h(g, s0);
                            else redo |= 1u << g;
                        }
                    } else {
                    GrpH s0, s1;
                    stage_a_h(0, s0);
#pragma unroll 1
                    for (int g = 0; g < NG - 2; g += 2) {
                        step_h(g, s0, s1);
                        step_h(g + 1, s1, s0);
                    }
                    step_h(NG - 2, s0, s1);
                    if (!__any_sync(0xffffffffu, s1.redo)) stage_b_h(NG - 1, s1);
                    else redo |= 1u << (NG - 1);
                    }
#pragma unroll 1
                    while (redo) {
                        const int g = __ffs((int)redo) - 1;
                        redo &= redo - 1;
                        Grp e;
                        stage_a(g, e);
                        stage_b(g, e, false);
                    }
                };

                // ---- fp32 screening path ---------------------------------------------------------------------------------
                // The lag of nearly every pair is decided in SINGLE precision, two pairs per instruction (FADD2 / FMUL2 /
                // FFMA2): B coordinates relative to their tile's origin, rounded to fp32 by the pre-pass (tile_f32_kernel), A
                // coordinates relative to the same origin, d2f = fma(dxf, dxf, dyf * dyf).  With R = the tile's half-extent
                // (<= f32_rcap) and u = 2^-24:  |d2f - d2| <= u (6 d2 + 5.66 R d)  (conversion of both points, the fp32
                // subtraction, product and fused sum; d2 = the reference's fp64 value, whose own rounding is 2^-29 of that).
                // The host turns that bound (x 1.05) into a band per threshold T: the table holds tmid = fl32(T) and hw with
                // [tmid - hw, tmid + hw] covering T +- bound(T).  A pair with d2f - tmid > hw is above T in the reference's
                // arithmetic, one with d2f - tmid < -hw is not; a group with a pair INSIDE a band (about 1 % of the warp-wide
                // groups on BASELINE config 2), like a group with a pair that may lie beyond its second lag, is redone by the
                // exact fp64 stages after the sweep — so the counts stay bit-exact.  Per pair: 2 packed fp32 instructions for
                // the distance instead of 4 fp64, 1 instead of 2 for the group's min / max, half a shared-memory wavefront less.
                struct GrpF {
                    float e[APT][U];           // d2f - tmid of the group's first lag
                    float hw[APT];
                    uint32_t off_lo[APT], off_hi[APT];
                    bool redo;
                };
                auto stage_a_f = [&](int g, GrpF &s) __attribute__((always_inline)) {
                    const ulonglong2 *btf = reinterpret_cast<const ulonglong2 *>(s_tf[F32S ? b : 0]) + g * (U / 2);
                    unsigned long long d[APT][U / 2];
                    float mn[APT], mx[APT];
#pragma unroll
                    for (int m = 0; m < U / 2; ++m) {
                        const ulonglong2 pb = btf[m];          // {x0, x1}, {y0, y1} of B points 2m, 2m + 1 of the group
#pragma unroll
                        for (int k = 0; k < APT; ++k) {
                            const unsigned long long dx = f2_sub(xa2[k], pb.x), dy = f2_sub(ya2[k], pb.y);
                            d[k][m] = f2_fma(dx, dx, f2_mul(dy, dy));
                            float lo, hi;
                            f2_unpack(d[k][m], lo, hi);
                            mn[k] = m ? fminf(fminf(mn[k], lo), hi) : fminf(lo, hi);
                            mx[k] = m ? fmaxf(fmaxf(mx[k], lo), hi) : fmaxf(lo, hi);
                        }
                    }
                    s.redo = false;
#pragma unroll
                    for (int k = 0; k < APT; ++k) {
#if PIB_F32_CLAMP
                        // clamp ahead of the shift: add + max fuse into one instruction
                        const int idx = min(max(__float_as_int(mn[k]) - base_f, 0), (lut_last << (LUT2_SHIFT + 3)) | 0xffff) >> (LUT2_SHIFT + 3);
#else
                        int idx = (__float_as_int(mn[k]) - base_f) >> (LUT2_SHIFT + 3);
                        idx = min(max(idx, 0), lut_last);
#endif
                        unsigned short q0;
                        asm("ld.shared.u16 %0, [%1];" : "=h"(q0) : "r"(lut_sa + 2u * (uint32_t)idx));
                        PIB_DBG_ASSERT((int)q0 < n_slots);
                        uint32_t ea = sloth_sa + 16u * (uint32_t)q0;
                        float tmid, hwq;
#if PIB_F32_LOOKUP2
                        if (!MULTI_FIX) {
                            // threshold and half-width of the table's slot first, the whole entry of the slot it decides on after
                            float tm, hw0;
                            asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(tm), "=f"(hw0) : "r"(ea));
                            ea += (__fsub_rn(mn[k], tm) > hw0) ? 16u : 0u;
                        }
#endif
                        asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(tmid), "=f"(hwq), "=r"(s.off_lo[k]), "=r"(s.off_hi[k]) : "r"(ea));
                        if (PIB_F32_LOOKUP2 && !MULTI_FIX) {
                        } else if (MULTI_FIX) {
                            while (__fsub_rn(mn[k], tmid) > hwq) {
                                ea += 16u;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(tmid), "=f"(hwq), "=r"(s.off_lo[k]), "=r"(s.off_hi[k]) : "r"(ea));
                            }
                        } else {
                            // the table cell is at most one slot low: the next entry is loaded by the few lanes that need it
                            asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tsub.rn.f32 t, %5, %0;\n\tsetp.gt.f32 p, t, %1;\n\t"
                                "@p ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+16];\n\t@p add.u32 %4, %4, 16;\n\t}"
                                : "+f"(tmid), "+f"(hwq), "+r"(s.off_lo[k]), "+r"(s.off_hi[k]), "+r"(ea) : "f"(mn[k]));
                        }
                        float cnext;
                        asm("ld.shared.f32 %0, [%1];" : "=f"(cnext) : "r"(cnext_sa + ((ea - sloth_sa) >> 2)));
                        s.hw[k] = hwq;
                        const unsigned long long nt2 = f2_pack(-tmid, -tmid);
                        float em = INFINITY;
#pragma unroll
                        for (int m = 0; m < U / 2; ++m) {
                            f2_unpack(f2_add(d[k][m], nt2), s.e[k][2 * m], s.e[k][2 * m + 1]);
                            em = fminf(fminf(em, fabsf(s.e[k][2 * m])), fabsf(s.e[k][2 * m + 1]));
                        }
                        // undecided pair, or a pair that may lie beyond the group's second lag
                        s.redo |= (em <= hwq) | (mx[k] >= cnext);
                    }
                };
                auto stage_b_f = [&](int g, const GrpF &s) __attribute__((always_inline)) {
                    const int j0 = g * U;
                    // second-lag sums of both A points from ONE pass over the group's values
                    double h1a[APT], h1b[APT], h2a[APT], h2b[APT];
                    int ch[APT];
#pragma unroll
                    for (int k = 0; k < APT; ++k) ch[k] = 0;
#pragma unroll
                    for (int u = 0; u < U; u += 2) {
                        double2 zz = *reinterpret_cast<const double2 *>(bz + j0 + u);
#pragma unroll
                        for (int k = 0; k < APT; ++k) {
#if PIB_F32_ZRELOAD
                            // a second load of the two values instead of register copies of their low words
                            if (k > 0) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(zz.x), "=d"(zz.y) : "r"(smem_u32(bz + j0 + u)));
#endif
                            int za_hi, zb_hi;
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(za_hi), "+r"(ch[k]) : "r"(__double2hiint(zz.x)), "f"(s.e[k][u]), "f"(s.hw[k]));
                            asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %3, %4;\n\tselp.b32 %0, %2, 0, p;\n\t@p add.s32 %1, %1, 1;\n\t}"
                                : "=r"(zb_hi), "+r"(ch[k]) : "r"(__double2hiint(zz.y)), "f"(s.e[k][u + 1]), "f"(s.hw[k]));
                            const double za = __hiloint2double(za_hi, __double2loint(zz.x));
                            const double zb = __hiloint2double(zb_hi, __double2loint(zz.y));
#if PIB_F32_CHAIN
                            // one chain per sum (the two A points give four independent chains): two fp64 additions fewer per
                            // A point and group than with even / odd partial sums
                            if (u == 0) {
                                h1a[k] = za + zb; h2a[k] = __fma_rn(zb, zb, __dmul_rn(za, za));
                            } else {
                                h1a[k] = (h1a[k] + za) + zb; h2a[k] = __fma_rn(zb, zb, __fma_rn(za, za, h2a[k]));
                            }
#else
                            if (u == 0) {
                                h1a[k] = za; h2a[k] = __dmul_rn(za, za);
                                h1b[k] = zb; h2b[k] = __dmul_rn(zb, zb);
                            } else {
                                h1a[k] += za; h2a[k] = __fma_rn(za, za, h2a[k]);
                                h1b[k] += zb; h2b[k] = __fma_rn(zb, zb, h2b[k]);
                            }
#endif
                        }
                    }
                    const double2 gs = bgs[g];
#pragma unroll
                    for (int k = 0; k < APT; ++k) {
#if PIB_F32_CHAIN
                        const double h1 = h1a[k], h2 = h2a[k];
                        (void)h1b; (void)h2b;
#else
                        const double h1 = h1a[k] + h1b[k], h2 = h2a[k] + h2b[k];
#endif
                        const uint32_t a_lo = acc_sa[k] + s.off_lo[k], a_hi = acc_sa[k] + s.off_hi[k];
                        const uint32_t n_lo = cnt_sa[k] + s.off_lo[k], n_hi = cnt_sa[k] + s.off_hi[k];
                        PIB_DBG_ASSERT(s.off_lo[k] <= (uint32_t)RING * SLOT_BYTES && s.off_hi[k] <= (uint32_t)RING * SLOT_BYTES &&
                                       s.off_lo[k] % SLOT_BYTES == 0 && s.off_hi[k] % SLOT_BYTES == 0);
                        double2 v_lo, v_hi;
                        uint32_t c_lo, c_hi;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v_lo.x), "=d"(v_lo.y) : "r"(a_lo) : "memory");
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v_hi.x), "=d"(v_hi.y) : "r"(a_hi) : "memory");
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(c_lo) : "r"(n_lo) : "memory");
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(c_hi) : "r"(n_hi) : "memory");
                        v_lo.x += gs.x - h1; v_lo.y += gs.y - h2;
                        v_hi.x += h1;        v_hi.y += h2;
                        c_lo = c_lo + (uint32_t)U - (uint32_t)ch[k];
                        c_hi = c_hi + (uint32_t)ch[k];
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a_lo), "d"(v_lo.x), "d"(v_lo.y) : "memory");
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a_hi), "d"(v_hi.x), "d"(v_hi.y) : "memory");
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(n_lo), "r"(c_lo) : "memory");
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(n_hi), "r"(c_hi) : "memory");
                    }
                };
                auto sweep_f32 = [&]() __attribute__((always_inline)) {
                    unsigned redo_f = 0;
#pragma unroll 1
                    for (int g = 0; g < NG; ++g) {
                        GrpF s0;
                        stage_a_f(g, s0);
                        if (!__any_sync(0xffffffffu, s0.redo)) stage_b_f(g, s0);
                        else redo_f |= 1u << g;
                    }
#pragma unroll 1
                    while (redo_f) {
                        const int g = __ffs((int)redo_f) - 1;
                        redo_f &= redo_f - 1;
                        Grp e;
                        stage_a(g, e);
                        stage_b(g, e, false);
                    }
                };

                auto sweep = [&](const bool wide_mode) __attribute__((always_inline)) {
                    if (diagonal || !p.group_path) {
                        // Per-pair path: tiles that overlap my CTA's own points (keep i < j only) and data whose groups of
                        // U points are too wide for the two-lag scheme (sparse points / fine lags / stretched ellipses:
                        // the host clears group_path).  Batches of 4 pairs: distances and table lookups of the batch are
                        // independent chains, the ring updates follow in order.
#pragma unroll 1
                        for (int j0 = 0; j0 < TILE; j0 += 4) {
                            long long d2[APT][4];
                            uint32_t ea[APT][4];
                            double zj[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const double2 pj = bxy[j0 + u];
                                zj[u] = bz[j0 + u];
#pragma unroll
                                for (int k = 0; k < APT; ++k) {
                                    d2[k][u] = __double_as_longlong(dist2(pj, k));
                                    int idx = ((int)(d2[k][u] >> 32) - base_hi) >> LUT2_SHIFT;
                                    idx = min(max(idx, 0), lut_last);
                                    unsigned short q0;
                                    asm("ld.shared.u16 %0, [%1];" : "=h"(q0) : "r"(lut_sa + 2u * (uint32_t)idx));
                                    ea[k][u] = slot_sa + 16u * (uint32_t)q0;
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
#pragma unroll
                                for (int k = 0; k < APT; ++k) {
                                    // exact slot: threshold steps on the full bit pattern (the table cell is at most a few slots low)
                                    long long T;
                                    uint32_t offs;
                                    while (true) {
                                        int tlo, thi, nx;
                                        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(tlo), "=r"(thi), "=r"(nx), "=r"(offs) : "r"(ea[k][u]));
                                        T = ((long long)thi << 32) | (unsigned)tlo;
                                        if (!(d2[k][u] > T)) break;
                                        ea[k][u] += 16u;
                                    }
                                    uint32_t off = offs & 0xffffu;
                                    if (!(jbase + j0 + u > ig[k])) off = TRASH16;                   // keep i < j only
                                    if (wide_mode && !((unsigned)((int)((ea[k][u] - slot_sa) >> 4) - wlo) < wlen)) off = TRASH16;
                                    double2 *ap = reinterpret_cast<double2 *>(ring + off * 16 + my_acc[k]);
                                    uint16_t *cp = reinterpret_cast<uint16_t *>(ring + off * 16 + my_cnt[k]);
                                    double2 a = *ap;
                                    a.x += zj[u]; a.y = __fma_rn(zj[u], zj[u], a.y);
                                    *ap = a;
                                    *cp = (uint16_t)(*cp + 1u);
                                }
                            }
                        }
                        return;
                    }
                    if (F32S && !wide_mode && tile_f32) {
                        sweep_f32();
                        return;
                    }
                    if (HIW && !wide_mode && p.hiw_ok && !f32_on) {
                        sweep_hiw();
                        return;
                    }
                    if (PIPE) {
                        Grp sa, sb;
                        stage_a(0, sa);
#pragma unroll 1
                        for (int g = 0; g < NG - 2; g += 2) {
                            stage_a(g + 1, sb);
                            stage_b(g, sa, wide_mode);
                            stage_a(g + 2, sa);
                            stage_b(g + 1, sb, wide_mode);
                        }
                        stage_a(NG - 1, sb);
                        stage_b(NG - 2, sa, wide_mode);
                        stage_b(NG - 1, sb, wide_mode);
                    } else {
#pragma unroll 2
                        for (int g = 0; g < NG; ++g) {
                            Grp s;
                            stage_a(g, s);
                            stage_b(g, s, wide_mode);
                        }
                    }
                };
                if (!wide) {
                    sweep(false);
                } else {
                    // this tile alone spans >= RING lags for my warp: sweep it once per window of RING lags, folding in
                    // between (no atomics, so the result stays deterministic)
                    fold();
                    for (int w0 = rng.x; w0 <= rng.y; w0 += RING) {
                        wlo = w0; wlen = (unsigned)(min(w0 + RING - 1, (int)rng.y) - w0 + 1);
                        sweep(true);
                        w_lo = w0;
                        fold();
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);          // my warp is done with buffer b
        }
        fold();                                               // my z_i changes with the next item
    }
    if (lane == 0 && evaluated) atomicAdd(p.pairs_evaluated, evaluated);
}

}  // namespace pib
