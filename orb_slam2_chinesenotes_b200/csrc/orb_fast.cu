// orb_fast.cu -- per-cell FAST-9 with NMS and the iniThFAST -> minThFAST retry, sm_100a.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:826-875
// (cv::FAST(cell, kps, iniThFAST, true), and again with minThFAST when that returns nothing).
//
// One block per STRIP: a run of up to ORB_FAST_STRIP horizontally adjacent 30-px cells of one
// cell row.  The kernel is bound by the integer/min-max issue rate (tools/ubench_pipes.cu), so
// it is organised around instructions per pixel:
//   * the strip is staged into shared memory as one image (evaluated rectangle + 3-px ring
//     apron), widened ONCE to 16 bits per pixel and stored TWICE: copy A holds pixel pairs that
//     start on an even column, copy B pairs that start on an odd column.  Each of the 16 ring
//     positions of a horizontally adjacent PIXEL PAIR is then one aligned 32-bit shared load of
//     a ready-made 16x2 operand -- no funnel shifts or byte permutes in the scoring loop;
//   * the score network works on the 16x2 operands with the full-rate 2-input half2 min/max
//     (VHMNMX; the values 0..255 are positive fp16 denormals, whose order is the integer
//     order), 47 operations per polarity (below) instead of the 80 of a sliding 3-input
//     network, with no data-dependent branch;
//   * a thread owns one pixel-pair column of the strip and walks down the rows, so index math
//     is paid per column, not per pixel;
//   * NMS reads a 16x2 score map with the same packing (3 rows x 3 words, column maxima,
//     two permutes for the left/right neighbours, one carry-free packed compare).
//
// Score (OpenCV cornerScore<16>, threshold independent), d[k] = v - ring[k]:
//     score = max( max_k min_{m<9} d[k+m], max_k min_{m<9} -d[k+m] ) - 1
// v is constant over the ring, so the network runs on the RAW ring values E[0..15]:
//     max_k min9(-d) = M1 - v,  M1 = max_k min_{m<9} E[k+m]
//     max_k min9( d) = v - M2,  M2 = min_k max_{m<9} E[k+m]
// Every 9-arc is an 8-arc starting at an EVEN position plus one of the two ring values next to
// it, and min/max distribute over each other, so with F[j] = min(E[2j..2j+7]):
//     M1 = max_j min( F[j], max(E[2j-1], E[2j+8]) )                     (indices mod 16)
// F comes from pair minima B[j] = min(E[2j],E[2j+1]) by doubling: 8 + 8 + 8 ops, then 8 + 8 + 7.
// A pixel is a FAST corner at threshold t iff score >= t.  NMS is the strict 3x3 maximum of the
// score map INSIDE the cell (FAST runs on the cropped cell image, so neighbours outside the
// cell's evaluated rectangle count as 0); both thresholds read the same map, hence the
// reference's retry is a per-cell choice of cut-off: iniThFAST if any NMS survivor of the cell
// reaches it, else minThFAST.
#include "orb_device.cuh"
#include "orb_launch.h"

#ifndef ORB_FAST_TPC
#define ORB_FAST_TPC 32      // threads per cell of the strip
#endif
#define FAST_NT (ORB_FAST_TPC * ORB_FAST_STRIP)
#ifndef ORB_FAST_FULLCOL
#define ORB_FAST_FULLCOL 64   // frames per launch from which a block walks a whole strip column
#endif

__device__ __forceinline__ uint32_t hmin2(const uint32_t a, const uint32_t b)
{
    uint32_t d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmax2(const uint32_t a, const uint32_t b)
{
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// The min/max network is bound by the ALU pipe (HMNMX2 issues at half the warp rate), while the FMA pipe idles.
// Where a min AND a max of the same two operands are needed, the pair is computed on the FMA pipe instead:
//     r = relu(a - b) (HFMA2.RELU: b * -1 + a),  max = b + r,  min = a - r            (HADD2)
// exact, because the operands are integers 0..255 held as fp16 denormals (multiples of 2^-24 below 2^-14:
// sums and differences are representable, nothing rounds, f16 arithmetic keeps subnormals).
#ifndef ORB_FAST_FMA
#define ORB_FAST_FMA 2      // bit 0: the 8 first-stage pairs, bit 1: the 8 lo/hi pairs on the FMA pipe; per 1024 frames: 0 5.24 ms, 1 5.03, 2 4.98, 3 5.09
#endif
#ifndef ORB_FAST_UNROLL
#define ORB_FAST_UNROLL 2
#endif
constexpr int kFastUnroll = ORB_FAST_UNROLL;
__device__ __forceinline__ void hminmax_fma(const uint32_t a, const uint32_t b, uint32_t& mn, uint32_t& mx)
{
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(0xbc00bc00u), "r"(a));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(mx) : "r"(b), "r"(r));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(mn) : "r"(a), "r"(r));
}
template <bool FMA>
__device__ __forceinline__ void hminmax(const uint32_t a, const uint32_t b, uint32_t& mn, uint32_t& mx)
{
    if (FMA) hminmax_fma(a, b, mn, mx);
    else { mn = hmin2(a, b); mx = hmax2(a, b); }
}

// packed (M1, M2) of a pixel pair from its 16 ring operands
__device__ __forceinline__ void fast_network(const uint32_t* E, uint32_t& M1, uint32_t& M2)
{
    uint32_t Bn[8], Bx[8], Qn[8], Qx[8], Y[8], Z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hminmax<(ORB_FAST_FMA & 1) != 0>(E[2 * j], E[2 * j + 1], Bn[j], Bx[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { Qn[j] = hmin2(Bn[j], Bn[(j + 1) & 7]); Qx[j] = hmax2(Bx[j], Bx[(j + 1) & 7]); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t Fn = hmin2(Qn[j], Qn[(j + 2) & 7]);                 // min E[2j .. 2j+7]
        const uint32_t Fx = hmax2(Qx[j], Qx[(j + 2) & 7]);                 // max E[2j .. 2j+7]
        uint32_t ln, lx;
        hminmax<(ORB_FAST_FMA & 2) != 0>(E[(2 * j + 15) & 15], E[(2 * j + 8) & 15], ln, lx);
        Y[j] = hmin2(Fn, lx);
        Z[j] = hmax2(Fx, ln);
    }
    M1 = hmax2(hmax2(hmax2(Y[0], Y[1]), hmax2(Y[2], Y[3])), hmax2(hmax2(Y[4], Y[5]), hmax2(Y[6], Y[7])));
    M2 = hmin2(hmin2(hmin2(Z[0], Z[1]), hmin2(Z[2], Z[3])), hmin2(hmin2(Z[4], Z[5]), hmin2(Z[6], Z[7])));
}

// n / d for 0 <= n < 2^20, 1 <= d < 2^10: (n + 0.5) / d is never within float rounding of an integer
__device__ __forceinline__ int fast_div(const int n, const int d)
{
    return __float2int_rz(__fmul_rn((float)n + 0.5f, __frcp_rn((float)d)));
}

// Which (level, strip column, cell rows) a block works on: filled per launch (rows_per_block follows the batch
// size: whole strip columns for big batches, single cell rows when few frames must fill the GPU).
struct FastGrid {
    int first[ORB_MAX_LEVELS + 1];   // first block of each level; [nlevels] = total
    int rows_per_block;
};

// Geometry of a strip's shared-memory images.  STATIC: compile-time strides for cells up to
// ORB_FAST_WC_STATIC wide (all offsets of the scoring loop become immediates); otherwise run-time.
//   tile row  = [copy A: WPC words][pad][copy B: WPC words][1] of the strip as ONE image (cells are contiguous,
//               so nothing is staged twice); copy A word j = strip columns (2j, 2j+1), copy B word j = columns
//               (2j+1, 2j+2); strip column 0 = first column of the first cell image.  Copy B starts OB words
//               into the row with OB = 1 (mod 32): a warp spans two cells, and when wCell is odd its second
//               cell reads the other copy -- this offset puts the two half-warps on disjoint banks.
//   score row = 1 + ncs * SP words, SP = np + 1: pixel pair p (evaluated columns 2p, 2p+1) of cell c at word
//               c*SP + p + 1; word c*SP is the zero apron between cells
// Pixel pair p of cell c has its ring window starting at strip column t = c*wCell + 2p.  For even t the odd
// ring offsets dx are words of copy A and the even ones words of copy B; for odd t the roles swap.  Either
// way the words are a[(3+dx)/2] (dx odd) and b[(2+dx)/2] (dx even) from two per-column base pointers.
// Everything that depends only on the strip column is set up once; the block then walks down its cell rows.
template <bool STATIC>
__device__ __forceinline__ void fast_strip_body(const OrbPlan& plan, const OrbBatch& io, uint32_t* smem, int* s_ctr, int* s_any,
                                                const int frame, const int l, const int sx, const int ci0, const int ci1)
{
    const OrbLevel& L = plan.lv[l];
    const int tid = threadIdx.x;
    const int cj0 = sx * ORB_FAST_STRIP;
    const int ncs = min(ORB_FAST_STRIP, L.ncx - cj0);
    const int wc = L.wCell, hc = L.hCell, maxBX = L.w - ORB_BORDER0, maxBY = L.h - ORB_BORDER0;
    const int np = (wc + 1) >> 1;                    // pixel pairs per cell row
    const int WPC = STATIC ? ORB_FAST_WPC_STATIC : orb_fast_wpc(ncs, wc);
    const int OB = orb_fast_ob(WPC), RS = OB + WPC + 1;         // even: rows stay 8-byte aligned
    const int SP = np + 1, SRS = ncs * SP + 1;
    uint32_t* tile = smem;                                            // (eh+6) x RS
    uint32_t* score = smem + plan.fast_tile_words;                    // (eh+2) x SRS
    uint32_t* surv = tile;                                            // NMS survivors (tile is dead by then)
    uint16_t* surv_tag = (uint16_t*)(score + plan.fast_score_words);  // cell | isA << 15
    uint32_t* raw = score + plan.fast_score_words + ((plan.fast_surv_max + 1) >> 1);   // (eh+6) x RW image words, as fetched
    const int RW = STATIC ? ORB_FAST_RW_STATIC : orb_fast_rw(ncs, wc);

    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    const int w = L.w;

    // ---- this thread's pixel-pair column: col = c * np + p, row group g of rgc
    const int ipr = ncs * np;
    const int rgc = max(fast_div(FAST_NT, ipr), 1);
    const int g = fast_div(tid, ipr), col = tid - g * ipr;
    int nv = 0, c = 0, xb = 0, aofs = 0, ab = 0, sofs = 0;
    if (g < rgc) {
        c = fast_div(col, np);
        const int p = col - c * np;
        const int X = (cj0 + c) * wc;                                  // border-frame x of the cell image's first column
        const int ew = min(wc, maxBX - 6 - (ORB_BORDER0 + X));         // evaluated width of this cell
        nv = min(max(ew - 2 * p, 0), 2);                               // valid pixels of the pair
        const int t = c * wc + 2 * p, odd = t & 1;
        aofs = (odd ? OB : 0) + (t >> 1);
        ab = (odd ? (t >> 1) + 1 : OB + (t >> 1)) - aofs;
        sofs = c * SP + p + 1;
        xb = X + 3 + 2 * p;
    }
    const uint32_t lanes = nv == 2 ? 0xffffffffu : 0x0000ffffu;
    // stored score = score - (minTh - 1) where score >= minTh, else 0 (monotone, so NMS is unchanged)
    const uint32_t bias = 0x00010001u * (uint32_t)(65536 - 257 - (plan.minTh - 1));
    const int iniBias = plan.iniTh - plan.minTh + 1;                   // stored score of a corner at iniThFAST
    // ---- this thread's staging column: 4-column group k, row group gs of rgs
    const int ncolS = min(WPC >> 1, ((ncs * wc + 7) >> 2) + 1);        // columns 0 .. ncs*wc+7 are read
    const int rgs = fast_div(FAST_NT, ncolS);
    const int gs = fast_div(tid, ncolS), ks = tid - gs * ncolS;
    const int x0 = ORB_BORDER0 + cj0 * wc;                             // level x of strip column 0
    const int dstep = rgs * RS;
    // ---- this thread's fetch column: image word jp of every rgp-th row.  Rows are fetched from the 4-byte
    // aligned address at or below their first pixel; bytes past the image width are zero-filled.
    const int rgp = fast_div(FAST_NT, RW);
    const int gp = fast_div(tid, RW), jp = tid - gp * RW;
    const bool fetch_full = x0 + 4 * jp + 8 <= w;

    auto prefetch = [&](const int ci) {
        const int y0 = ORB_BORDER0 + ci * hc, th = min(y0 + hc + 6, maxBY) - y0;
        if (gp < rgp) {
            const uint8_t* rowp = src + (size_t)(y0 + gp) * pitch + x0;
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(raw + gp * RW + jp);
            if ((pitch & 3) == 0) {
                // every row starts at the same offset inside its word (all pyramid levels, and level 0 when the caller's
                // pitch is a multiple of 4): source word and byte count are set up once, the loop only steps
                const uintptr_t a = (uintptr_t)rowp;
                const int s = (int)(a & 3);
                const uint8_t* q = (const uint8_t*)(a - s) + 4 * jp;
                const int n = fetch_full ? 4 : min(max(w - (x0 - s + 4 * jp), 0), 4);
                const size_t qstep = (size_t)rgp * pitch;
                const uint32_t dstep4 = 4u * rgp * RW;
                for (int r = gp; r < th; r += rgp, q += qstep, dst += dstep4)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(q), "r"(n) : "memory");
            } else {
                for (int r = gp; r < th; r += rgp, rowp += (size_t)rgp * pitch, dst += 4u * rgp * RW) {
                    const uintptr_t a = (uintptr_t)rowp;
                    const int s = (int)(a & 3);
                    const uint8_t* q = (const uint8_t*)(a - s) + 4 * jp;
                    int n = 4;
                    if (!fetch_full) n = min(max(w - (x0 - s + 4 * jp), 0), 4);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(q), "r"(n) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(ci0);

    for (int ci = ci0; ci < ci1; ++ci) {
        const int y0 = ORB_BORDER0 + ci * hc, y1 = min(y0 + hc + 6, maxBY);
        const int eh = y1 - y0 - 6;                  // evaluated rows y0+3 .. y1-4
        if (eh <= 0) break;
        const int th = eh + 6;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- staging: a thread owns a 4-column group and walks down the rows of the fetched image words
        if (gs < rgs) {
            uint32_t* dA = tile + 2 * ks + gs * RS;
            const uint32_t* rq = raw + gs * RW + ks;
            uint32_t s = (uint32_t)((uintptr_t)(src + (size_t)(y0 + gs) * pitch + x0) & 3);
            const uint32_t sinc = (uint32_t)(rgs * pitch) & 3u;
#pragma unroll 4
            for (int r = gs; r < th; r += rgs, rq += rgs * RW, dA += dstep, s = (s + sinc) & 3u) {
                const uint32_t q0 = rq[0], q1 = rq[1];
                const uint32_t w0 = __funnelshift_r(q0, q1, 8 * s);          // columns 4k .. 4k+3
                const uint32_t c4 = __byte_perm(q1, 0, 0x4440u | s);         // column 4k+4
                *(uint2*)dA = make_uint2(__byte_perm(w0, 0, 0x4140), __byte_perm(w0, 0, 0x4342));
                dA[OB] = __byte_perm(w0, 0, 0x4241);
                dA[OB + 1] = __byte_perm(w0, c4, 0x5453);
            }
        }
        {   // score map = 0 (aprons must be; interior is only written where non-zero)
            uint4* z = (uint4*)score;
            const int n4 = ((eh + 2) * SRS + 3) >> 2;
            for (int i = tid; i < n4; i += FAST_NT) z[i] = make_uint4(0, 0, 0, 0);
        }
        if (tid < ORB_FAST_STRIP) s_any[tid] = 0;
        if (tid < 4) s_ctr[tid] = 0;
        __syncthreads();
        if (ci + 1 < ci1) prefetch(ci + 1);                           // overlaps the scoring of this cell row

        // ---- a thread owns a pixel-pair column of the strip and a contiguous run of rows [ya, yb)
        const int chunk = fast_div(eh + rgc - 1, rgc);
        const int ya = g * chunk, yb = min(eh, ya + chunk);
        if (nv) {   // scores
            const uint32_t* a = tile + aofs + ya * RS;                    // ring row dy = -3 of evaluated row ya
            uint32_t* sc = score + sofs + (ya + 1) * SRS;
#pragma unroll (kFastUnroll)
            for (int ly = ya; ly < yb; ++ly, a += RS, sc += SRS) {
                const uint32_t* b = a + ab;
                uint32_t E[16];
                E[0] = b[6 * RS + 1];    //  ( 0, 3)
                E[1] = a[6 * RS + 2];    //  ( 1, 3)
                E[2] = b[5 * RS + 2];    //  ( 2, 2)
                E[3] = a[4 * RS + 3];    //  ( 3, 1)
                E[4] = a[3 * RS + 3];    //  ( 3, 0)
                E[5] = a[2 * RS + 3];    //  ( 3,-1)
                E[6] = b[1 * RS + 2];    //  ( 2,-2)
                E[7] = a[0 * RS + 2];    //  ( 1,-3)
                E[8] = b[0 * RS + 1];    //  ( 0,-3)
                E[9] = a[0 * RS + 1];    //  (-1,-3)
                E[10] = b[1 * RS + 0];   //  (-2,-2)
                E[11] = a[2 * RS + 0];   //  (-3,-1)
                E[12] = a[3 * RS + 0];   //  (-3, 0)
                E[13] = a[4 * RS + 0];   //  (-3, 1)
                E[14] = b[5 * RS + 0];   //  (-2, 2)
                E[15] = a[6 * RS + 1];   //  (-1, 3)
                const uint32_t v = b[3 * RS + 1];
                uint32_t M1, M2;
                fast_network(E, M1, M2);
                // per 16-bit lane, carry free: bright + 256 = M1 + (256 - v), dark + 256 = (v + 256) - M2
                const uint32_t br = M1 + (0x01000100u - v), dk = (v + 0x01000100u) - M2;
                const uint32_t t = hmax2(br, dk);                                   // score + 257
                const uint32_t out = __viaddmax_s16x2_relu(t, bias, 0u) & lanes;    // max(score - minTh + 1, 0)
                if (out) *sc = out;
            }
        }
        __syncthreads();

        // ---- strict 3x3 maximum inside each cell: rows slide through registers, 3 loads per pixel pair.
        // The loop only records WHICH lanes survive (2 bits per row); the rare survivors are pushed afterwards.
        if (nv) {
            const uint32_t* sc = score + sofs + (ya + 1) * SRS;           // row of evaluated row ya
            uint32_t u0 = sc[-SRS - 1], u1 = sc[-SRS], u2 = sc[-SRS + 1];
            uint32_t m0 = sc[-1], m1 = sc[0], m2 = sc[1];
            for (int yq = ya; yq < yb; yq += 32) {                        // 2 bits per row in a 64-bit mask
                const int ye = min(yb, yq + 32);
                unsigned long long kept = 0;
#pragma unroll 3
                for (int ly = yq; ly < ye; ++ly, sc += SRS) {
                    const uint32_t d0 = sc[SRS - 1], d1 = sc[SRS], d2 = sc[SRS + 1];
                    const uint32_t Lw = __vimax3_s16x2(u0, m0, d0), Rw = __vimax3_s16x2(u2, m2, d2);
                    const uint32_t Uw = hmax2(u1, d1), Cw = hmax2(Uw, m1);
                    // neighbours of (x | x+1): columns (x-1 | x) and (x+1 | x+2) over three rows, own column above/below
                    const uint32_t nb = __vimax3_s16x2(__byte_perm(Lw, Cw, 0x5432), __byte_perm(Cw, Rw, 0x5432), Uw);
                    // lane > neighbour maximum  <=>  bit 15 of (lane + 32768 - nb - 1); lanes stay in 0..65535, no borrow
                    const uint32_t keep = ((m1 | 0x80008000u) - nb - 0x00010001u) & 0x80008000u;
                    kept = (kept << 2) | ((keep >> 15) & 1u) | (keep >> 30);
                    u0 = m0; u1 = m1; u2 = m2; m0 = d0; m1 = d1; m2 = d2;
                }
                // kept: row ye-1 in bits 0..1, row ye-2 in bits 2..3, ...; sc is now the row of evaluated row ye
                while (kept) {
                    const int bit = __ffsll((long long)kept) - 1;
                    kept &= kept - 1;
                    const int back = bit >> 1, hi = bit & 1;
                    const int r = (int)((sc[-(back + 1) * SRS] >> (16 * hi)) & 0xffffu);
                    const int slot = atomicAdd(&s_ctr[0], 1);
                    // border-frame coordinates (src/ORBextractor.cc:868-869): cell-local + (j*wCell, i*hCell)
                    surv[slot] = orb_pack(xb + hi, ye - 1 - back + 3 + ci * hc, r + plan.minTh - 1);
                    const bool isA = r >= iniBias;
                    surv_tag[slot] = (uint16_t)(c | (isA ? 0x8000 : 0));
                    if (isA) s_any[c] = 1;
                }
            }
        }
        __syncthreads();
        // ---- per-cell cut-off: keep the iniThFAST survivors, or all of them if the cell has none (:857-861)
        const int nsurv = s_ctr[0];
        int mykeep = 0;
        for (int j = tid; j < nsurv; j += FAST_NT) {
            const int t = surv_tag[j];
            if ((t & 0x8000) || !s_any[t & 0x7fff]) ++mykeep;
        }
        if (mykeep) atomicAdd(&s_ctr[1], mykeep);
        __syncthreads();
        const int nkeep = s_ctr[1];
        if (nkeep) {
            if (tid == 0) s_ctr[3] = atomicAdd(&io.cand_count[frame * ORB_MAX_LEVELS + l], nkeep);
            __syncthreads();
            const int base = s_ctr[3];
            uint32_t* out = io.cand + (size_t)frame * plan.cand_per_frame + L.cand_off;
            for (int j = tid; j < nsurv; j += FAST_NT) {
                const int t = surv_tag[j];
                if ((t & 0x8000) || !s_any[t & 0x7fff]) {
                    const int slot = base + atomicAdd(&s_ctr[2], 1);
                    if (slot < L.cand_cap) out[slot] = surv[j];
                }
            }
        }
        __syncthreads();                                              // the next cell row reuses everything
    }
}

__global__ void __launch_bounds__(FAST_NT) k_fast_strips(const __grid_constant__ OrbPlan plan, const OrbBatch io, const __grid_constant__ FastGrid fg)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_ctr[4], s_any[ORB_FAST_STRIP];
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= fg.first[l + 1]) ++l;
    const int idx = blockIdx.x - fg.first[l];
    const int spr = plan.lv[l].spr;
    const int rb = fast_div(idx, spr), sx = idx - rb * spr;
    const int ci0 = rb * fg.rows_per_block, ci1 = min(plan.lv[l].ncy, ci0 + fg.rows_per_block);
    if (plan.lv[l].wCell <= ORB_FAST_WC_STATIC)
        fast_strip_body<true>(plan, io, smem, s_ctr, s_any, blockIdx.y, l, sx, ci0, ci1);
    else
        fast_strip_body<false>(plan, io, smem, s_ctr, s_any, blockIdx.y, l, sx, ci0, ci1);
}

size_t orb_fast_smem_bytes(const OrbPlan& plan)
{
    return ((size_t)plan.fast_tile_words + plan.fast_score_words + ((plan.fast_surv_max + 1) >> 1) + plan.fast_raw_words) * 4 + 32;
}

cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    if (plan.total_strips == 0) return cudaSuccess;
    const size_t smem = orb_fast_smem_bytes(plan);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, ORB_SMEM_OPTIN);   // a constant: the attribute is per-function state shared by all host threads (a per-launch value races)
        if (e != cudaSuccess) return e;
    }
    // Cell rows per block: whole strip columns once the batch alone fills the GPU several times over (the
    // per-column setup is then paid once per column), single cell rows for a few frames (latency).
    int maxrows = 1;
    for (int l = 0; l < plan.nlevels; ++l) if (plan.lv[l].ncy > maxrows) maxrows = plan.lv[l].ncy;
    FastGrid fg;
    fg.rows_per_block = batch >= ORB_FAST_FULLCOL ? maxrows : batch >= 16 ? 4 : 1;
    int n = 0;
    for (int l = 0; l < plan.nlevels; ++l) {
        fg.first[l] = n;
        n += plan.lv[l].spr * ((plan.lv[l].ncy + fg.rows_per_block - 1) / fg.rows_per_block);
    }
    for (int l = plan.nlevels; l <= ORB_MAX_LEVELS; ++l) fg.first[l] = n;
    if (n == 0) return cudaSuccess;
    k_fast_strips<<<dim3(n, batch), FAST_NT, smem, st>>>(plan, io, fg);
    return cudaGetLastError();
}
