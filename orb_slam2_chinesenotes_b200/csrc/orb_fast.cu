// orb_fast.cu -- per-cell FAST-9 with NMS and the iniThFAST -> minThFAST retry, sm_100a.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:826-875
// (cv::FAST(cell, kps, iniThFAST, true), and again with minThFAST when that returns nothing).
//
// Measured facts the design rests on (tools/ubench_fast.cu, profiles/r02_ubench_fast.txt): HMNMX2, the 3-input
// VHMNMX, PRMT, LOP3, IADD3 (ALU pipe) and HADD2 / HFMA2 (FMA pipe) all issue at ~1.9 warp instructions per clock
// and SM, a stream that mixes the two pipes tops out at ~2.9, so the kernel is bound by INSTRUCTIONS, mostly those
// of the ALU pipe.  Everything here is organised to issue fewer of them per pixel:
//
//   * ONE WARP = ONE BAND: up to 64 evaluated columns (one or two 30-px cells) walked down their cell rows by a
//     warp that shares nothing with any other warp -- no block barrier anywhere, only __syncwarp and the warp's
//     own mbarrier.  A lane owns the pixel pair (2p, 2p+1) of the band, whatever cell each pixel lies in.
//   * DATA PATH: the TMA unit.  Where the level's layout is 16-byte aligned (every pyramid level; level 0 when the
//     caller's base, pitch and frame stride allow) ONE tensor-map copy (cp.async.bulk.tensor.3d: x, y, frame)
//     brings the whole chunk, starting at the 16-byte aligned column at or below its first pixel, and the warp's
//     mbarrier counts the bytes; otherwise (level 0 of a caller with an odd pitch, where every row starts at another
//     offset inside its 16 bytes, so that no TMA form fits) lane j copies the aligned word j of every row with
//     cp.async.  Either way the copy of chunk k+1 is in flight while chunk k is scored.
//   * WIDENING on the LSU and FMA pipes: the raw bytes are re-read one by one (LDS.U8 has no alignment rule, so
//     no funnel shifts) and paired with one IMAD each into the two 16-bit copies of the tile -- copy A holds the
//     pixel pairs that start on an even tile column, copy B those that start on an odd one -- so every ring
//     operand of a pixel pair is one aligned, bank-conflict-free 32-bit load (consecutive lanes, consecutive words).
//   * the score network runs on the 16x2 operands with half2 min/max (the values 0..255 are positive fp16
//     denormals, whose order is the integer order): 36 operations per polarity, the 3-input ones fused by ptxas
//     into VHMNMX; min/max PAIRS of the same two operands can be moved to the idle FMA pipe (ORB_FAST_FMA);
//   * the score is finished on the FMA pipe in signed fp16 arithmetic (exact: all values are integers below 2048
//     in units of 2^-24), masked and biased by one HFMA2.RELU;
//   * NMS needs no score map: the scores of the last three rows of the lane's column pair live in registers, the
//     neighbouring columns come from two shuffles, cell and band borders are folded into the two PRMT selectors;
//   * a survivor is RECORDED, not handled, in the row loop: a column of a row pair holds at most one strict 3x3
//     maximum, so one word per row pair and lane (per 16-bit half: the stored score, times 256 if it lies in the
//     odd row) goes to a per-warp stash with one IMAD and one store; survivor counts per column accumulate on
//     the FMA pipe as fp16 numbers;
//   * per cell row the warp decides the cut-off of each cell (below), adds its total to the frame's candidate
//     counter with ONE atomic and lets every lane write its own survivors.
//
// Score (OpenCV cornerScore<16>, threshold independent), d[k] = v - ring[k]:
//     score = max( max_k min_{m<9} d[k+m], max_k min_{m<9} -d[k+m] ) - 1
// v is constant over the ring, so the network runs on the RAW ring values E[0..15]:
//     max_k min9(-d) = M1 - v,  M1 = max_k min_{m<9} E[k+m]
//     max_k min9( d) = v - M2,  M2 = min_k max_{m<9} E[k+m]
// Every 9-arc is an 8-arc starting at an EVEN position plus one of the two ring values next to
// it, and min/max distribute over each other, so with F[j] = min(E[2j..2j+7]):
//     M1 = max_j min( F[j], max(E[2j-1], E[2j+8]) )                     (indices mod 16)
// A pixel is a FAST corner at threshold t iff score >= t.  NMS is the strict 3x3 maximum of the
// score map INSIDE the cell (FAST runs on the cropped cell image, so neighbours outside the
// cell's evaluated rectangle count as 0); both thresholds read the same map, hence the
// reference's retry is a per-cell choice of cut-off: iniThFAST if any NMS survivor of the cell
// reaches it, else minThFAST.
#include "orb_device.cuh"
#include "orb_launch.h"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>

#ifndef ORB_FAST_R
#define ORB_FAST_R 16         // evaluated rows per chunk (even); the tile holds R + 6 rows (per 512 frames: 10 rows 2.12 ms, 12 1.99, 14 2.02, 16 1.99, 20 2.06, 24 2.14)
#endif
#ifndef ORB_FAST_WPB
#define ORB_FAST_WPB 1        // warps (bands) per block; the warps of a block share nothing (per 512 frames: 1 warp 1.99 ms, 2 warps 2.07, 4 warps 2.06)
#endif
#ifndef ORB_FAST_FMA
#define ORB_FAST_FMA 0x00ffu  // bit j: first-stage pair j, bit 8+j: lo/hi pair j computed on the FMA pipe
#endif
#ifndef ORB_FAST_FULLCOL
#define ORB_FAST_FULLCOL 256  // frames per launch from which a warp walks a whole band column
#endif
#ifndef ORB_FAST_UNROLL
#define ORB_FAST_UNROLL 2     // row pairs per trip of the scoring loop (per 512 frames: 1 pair 1.963 ms, 2 pairs 1.931)
#endif
#ifndef ORB_FAST_MIDROWS
#define ORB_FAST_MIDROWS 2    // cell rows per warp for 16 .. ORB_FAST_FULLCOL-1 frames (64 frames: 2 rows 0.27 ms, 4 rows 0.29, whole columns 0.35)
#endif
#ifndef ORB_FAST_MINBLK
#define ORB_FAST_MINBLK 16    // resident blocks per SM the register allocation must allow
#endif

constexpr int kFastUnroll = ORB_FAST_UNROLL;
constexpr int kR = ORB_FAST_R;
constexpr int kTileRows = kR + 6;
constexpr int kCopyW = 36;                 // 32-bit words per copy of a tile row (72 pixels)
constexpr int kRS = 2 * kCopyW;            // tile row: [copy A][copy B]
constexpr int kRawPitch = 96;              // bytes per raw row, row-by-row copies: 15 (alignment) + 70 pixels, rounded up to 16
constexpr int kBoxW = kRawPitch;           // tensor-map copies: the box starts at the 16-byte aligned column at or below the first pixel (measured: the TMA
                                           // unit raises an illegal-instruction fault on an innermost coordinate that is not a multiple of 16 bytes)
constexpr int kNT = 32 * ORB_FAST_WPB;
static_assert(kR % 2 == 0 && kR >= 2, "chunks hold whole row pairs");
static_assert(kTileRows <= 32, "one lane per tile row issues the bulk copy");

__device__ __forceinline__ uint32_t hmin2(const uint32_t a, const uint32_t b) { uint32_t d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmax2(const uint32_t a, const uint32_t b) { uint32_t d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hadd2(const uint32_t a, const uint32_t b) { uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hsub2(const uint32_t a, const uint32_t b) { uint32_t d; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmul2(const uint32_t a, const uint32_t b) { uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hfma2(const uint32_t a, const uint32_t b, const uint32_t c) { uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hfma2_relu(const uint32_t a, const uint32_t b, const uint32_t c) { uint32_t d; asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hsetgt2(const uint32_t a, const uint32_t b) { uint32_t d; asm("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }   // 1.0 / 0.0 per lane
__device__ __forceinline__ uint32_t hsetge2(const uint32_t a, const uint32_t b) { uint32_t d; asm("set.ge.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

// Where a min AND a max of the same two operands are needed, the pair can be computed on the FMA pipe:
//     r = relu(a - b) (HFMA2.RELU: b * -1 + a),  max = b + r,  min = a - r            (HADD2)
// exact, because the operands are integers 0..255 held as fp16 denormals (multiples of 2^-24 below 2^-14:
// sums and differences are representable, nothing rounds, f16 arithmetic keeps subnormals).
template <bool FMA>
__device__ __forceinline__ void hminmax(const uint32_t a, const uint32_t b, uint32_t& mn, uint32_t& mx)
{
    if (FMA) {
        const uint32_t r = hfma2_relu(b, 0xbc00bc00u, a);
        mx = hadd2(b, r);
        mn = hsub2(a, r);
    } else {
        mn = hmin2(a, b);
        mx = hmax2(a, b);
    }
}

// packed (M1, M2) of a pixel pair from its 16 ring operands
__device__ __forceinline__ void fast_network(const uint32_t* E, uint32_t& M1, uint32_t& M2)
{
    uint32_t Bn[8], Bx[8], Qn[8], Qx[8], Y[8], Z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if ((ORB_FAST_FMA >> j) & 1u) hminmax<true>(E[2 * j], E[2 * j + 1], Bn[j], Bx[j]);
        else hminmax<false>(E[2 * j], E[2 * j + 1], Bn[j], Bx[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { Qn[j] = hmin2(Bn[j], Bn[(j + 1) & 7]); Qx[j] = hmax2(Bx[j], Bx[(j + 1) & 7]); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t Fn = hmin2(Qn[j], Qn[(j + 2) & 7]);                 // min E[2j .. 2j+7]
        const uint32_t Fx = hmax2(Qx[j], Qx[(j + 2) & 7]);                 // max E[2j .. 2j+7]
        uint32_t ln, lx;
        if ((ORB_FAST_FMA >> (8 + j)) & 1u) hminmax<true>(E[(2 * j + 15) & 15], E[(2 * j + 8) & 15], ln, lx);
        else hminmax<false>(E[(2 * j + 15) & 15], E[(2 * j + 8) & 15], ln, lx);
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

// ---- mbarrier + bulk copy (one barrier per warp, arrival count 1)
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tensor_g2s(const uint32_t dst, const void* map, const int x, const int y, const int z, const uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

struct alignas(64) FastMapsArg { unsigned char map[ORB_MAX_LEVELS][128]; };

// Which (level, band group, cell rows) a block works on: filled per launch (rows_per_block follows the batch
// size: whole band columns for big batches, single cell rows when few frames must fill the GPU).
struct FastGrid {
    int first[ORB_MAX_LEVELS + 1];   // first block of each level; [nlevels] = total
    int rows_per_block;
    unsigned tensor_levels;          // bit l: level l is fetched through its tensor map
};

// Per-warp shared memory: [raw: kTileRows x kRawPitch bytes, 128-byte aligned (tensor-map copies demand it)][tile:
// kTileRows x kRS words][stash: (slots + 1) x 32 words][mbarrier].
//   tile row  = [copy A: 36 words][copy B: 36 words]; tile column u = band column t + 3 (the 3-px ring apron on the
//               left); copy A word j = tile columns (2j, 2j+1), copy B word j = columns (2j+1, 2j+2).  Lane p
//               scores band columns (2p, 2p+1), i.e. tile columns (2p+3, 2p+4): ring offsets dx = -2, 0, +2 are
//               words p, p+1, p+2 of copy B and dx = -3, -1, +1, +3 words p .. p+3 of copy A.
//   stash     = one word per (row pair, lane): per 16-bit half (column) 0, or the stored score of the surviving
//               pixel of that column, times 256 if it lies in the odd row of the pair.  The two pixels of a column
//               are neighbours, so a half never holds two survivors; the two halves of a lane can both hold one
//               only where the lane straddles two cells (odd cell width).
__host__ __device__ inline int fast_warp_bytes(const int stash_slots)
{
    return (((kTileRows * kRawPitch + 127) & ~127) + kTileRows * kRS * 4 + (stash_slots + 1) * 128 + 16 + 127) & ~127;
}

__global__ void __launch_bounds__(kNT, ORB_FAST_MINBLK) k_fast_bands(const __grid_constant__ OrbPlan plan, const OrbBatch io, const __grid_constant__ FastGrid fg,
                                                                      const __grid_constant__ FastMapsArg maps)
{
    extern __shared__ __align__(128) unsigned char fast_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= fg.first[l + 1]) ++l;
    const OrbLevel& L = plan.lv[l];
    const int idx = blockIdx.x - fg.first[l];
    const int bpr = (L.fbands + ORB_FAST_WPB - 1) / ORB_FAST_WPB;      // blocks per group of cell rows
    const int rb = fast_div(idx, bpr);
    const int band = (idx - rb * bpr) * ORB_FAST_WPB + warp;
    if (band >= L.fbands) return;                                      // warps share nothing: no barrier follows
    const int frame = blockIdx.y;
    const int ci0 = rb * fg.rows_per_block, ci1 = min(L.ncy, ci0 + fg.rows_per_block);

    unsigned char* wsm = fast_smem + (size_t)warp * fast_warp_bytes(plan.fast_stash_slots);
    unsigned char* raw = wsm;
    uint32_t* tile = (uint32_t*)(wsm + ((kTileRows * kRawPitch + 127) & ~127));
    uint32_t* stash = tile + kTileRows * kRS + lane;                           // this lane's column of the stash; slot 0 is a dummy
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(tile + kTileRows * kRS + (plan.fast_stash_slots + 1) * 32);
    const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(raw);

    // ---- band geometry (src/ORBextractor.cc:826-848): cells cj0 .. cj0+ncb-1 of every cell row
    const int wc = L.wCell, hc = L.hCell, maxBX = L.w - ORB_BORDER0, maxBY = L.h - ORB_BORDER0;
    const int cpb = L.fcpb;                                            // cells per band: 2 while two cells fit 64 columns
    const int cj0 = band * cpb, ncb = min(cpb, L.ncx - cj0);
    const int X0 = cj0 * wc;                                           // border-frame x of the band's first cell image
    // evaluated columns of the band: cells are contiguous, only the last cell of a level can be narrower
    const int tw = min(ncb * wc, maxBX - 6 - (ORB_BORDER0 + X0));
    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    src += ORB_BORDER0 + X0;                                           // level address of tile column 0, row 0
    const bool tensor = (fg.tensor_levels >> l) & 1u;                  // one tensor-map copy per chunk; else one bulk copy per row
    const void* tmap = maps.map[l];
    const int tx0 = (ORB_BORDER0 + X0) & ~15;                          // first column of the tensor box

    // ---- this lane's pixel pair: band columns t0 = 2 * lane (lo half) and t0 + 1 (hi half)
    const int t0 = 2 * lane;
    const bool vlo = t0 < tw, vhi = t0 + 1 < tw;
    const int clo = (cpb > 1 && t0 >= wc) ? 1 : 0, chi = (cpb > 1 && t0 + 1 >= wc) ? 1 : 0;   // cell of each half
    const uint32_t lm = (vlo ? 0x00003c00u : 0u) | (vhi ? 0x3c000000u : 0u);   // 1.0 in the valid halves
    // NMS neighbours: (left of lo | left of hi) = (Lw.hi | Cw.lo), (right of lo | right of hi) = (Cw.hi | Rw.lo);
    // a neighbour in another cell, outside the band or in another warp reads byte 1 of the first operand, the
    // high byte of a score below 256, i.e. zero
    const bool lo_first = t0 == 0 || (cpb > 1 && t0 == wc);
    const bool hi_first = cpb > 1 && t0 + 1 == wc;
    const bool lo_last = (cpb > 1 && t0 == wc - 1) || t0 == tw - 1;
    const bool hi_last = (cpb > 1 && t0 + 1 == wc - 1) || t0 + 1 >= tw - 1 || lane == 31;
    const uint32_t selL = (lo_first ? 0x0011u : 0x0032u) | (hi_first ? 0x1100u : 0x5400u);
    const uint32_t selR = (lo_last ? 0x0011u : 0x0032u) | (hi_last ? 0x1100u : 0x5400u);
    const uint32_t bias = 0x80008000u | (0x00010001u * (uint32_t)plan.minTh);            // -minTh * 2^-24 per half
    const int iniBias = plan.iniTh - plan.minTh + 1;                                      // stored score of a corner at iniThFAST
    const uint32_t iniV = 0x00010001u * (uint32_t)min(max(iniBias, 1), 1023);          // every survivor is >= 1
    const uint32_t* po = tile + lane;                // odd ring offsets: words po[0..3] = dx -3, -1, +1, +3
                                                     // even ring offsets: words po[kCopyW .. kCopyW+2] = dx -2, 0, +2
    // widening: lane = (sub-row, 8-column group); 27 lanes work on 3 tile rows per pass
    const int wsub = lane / 9, wk = lane - 9 * wsub;
    const uint32_t pitch15 = tensor ? 0u : (uint32_t)pitch & 3u;   // how the offset of a row inside its aligned word moves from row to row

    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // Issue the bulk copies of tile rows [0, trows) whose first image row is y: lane r fetches row r from the
    // 16-byte aligned address at or below its first pixel; the bytes stay inside the image row (the band starts
    // at column >= 16 and its last tile column lies >= 16 pixels before the row's end).
    auto fetch = [&](const int y, const int trows) {
        if (tensor) {
            // the whole box (kTileRows rows, zero filled outside the image) always arrives
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the raw rows were read through the generic proxy
                mbar_expect_tx(bar, kBoxW * kTileRows);
                tensor_g2s(raw_s, tmap, tx0, y, frame, bar);
            }
            return;
        }
        // Layouts that are not 16-byte aligned (level 0 of a caller with an odd pitch): no TMA form fits, because every
        // row starts at a different offset inside its 16 bytes.  Lane j copies the aligned 32-bit word j of every row
        // with cp.async (19 words cover the 3 + 70 + 3 bytes a row can need); the widening reads from byte (row & 3).
        // (One bulk copy per row, issued by lane r, was the first version: ptxas serialises the 22 per-lane UBLKCP
        // through an ELECT loop of ~10 instructions each, 7 % of the kernel's instructions.)
        if (lane < 19) {
            const uint8_t* rowp = src + (size_t)y * pitch;
            uint32_t dst = raw_s + 4u * lane;
            for (int r = 0; r < trows; ++r, rowp += pitch, dst += kRawPitch) {
                const uint8_t* q = (const uint8_t*)((uintptr_t)rowp & ~(uintptr_t)3) + 4 * lane;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(q) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // chunk schedule of a cell row: evaluated rows [ya, yb) of the cell; returns false when the cell row is empty
    uint32_t parity = 0;
    {
        const int y0 = ORB_BORDER0 + ci0 * hc, eh = min(y0 + hc + 6, maxBY) - y0 - 6;
        if (ci0 < ci1 && eh > 0) fetch(y0, min(eh, kR) + 6);
    }

    for (int ci = ci0; ci < ci1; ++ci) {
        const int y0 = ORB_BORDER0 + ci * hc;
        const int eh = min(y0 + hc + 6, maxBY) - y0 - 6;               // evaluated rows y0+3 .. y0+3+eh-1
        if (eh <= 0) break;
        uint32_t u = 0, m = 0, carry = 0;                              // scores of the two rows above; even-row survivors
        uint32_t cntAll = 0, cntA = 0;                                 // survivors / survivors at iniThFAST per column (fp16 counts)
        uint32_t smask = 0;                                            // one bit per record of this cell row (newest in bit 0): the slot holds a survivor

        for (int ya = 0; ya < eh; ya += kR) {
            const int nr = min(eh - ya, kR), trows = nr + 6;
            if (tensor) { mbar_wait(bar, parity); parity ^= 1u; }
            else { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(); }
            // ---- widen: raw bytes -> copies A and B
            {
                const uint32_t s0 = tensor ? (uint32_t)(ORB_BORDER0 + X0 - tx0) : (uint32_t)((uintptr_t)(src + (size_t)(y0 + ya) * pitch) & 3u);
                if (wsub < 3) {
                    auto widen_row = [&](const unsigned char* rq, uint32_t* dst) {
                        uint32_t b[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) b[i] = rq[i];             // LDS.U8: no alignment rule, no funnel shift
                        uint32_t A[4], B[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(A[i]) : "r"(b[2 * i + 1]), "r"(b[2 * i]));
                            asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(B[i]) : "r"(b[2 * i + 2]), "r"(b[2 * i + 1]));
                        }
                        *(uint4*)dst = make_uint4(A[0], A[1], A[2], A[3]);
                        *(uint4*)(dst + kCopyW) = make_uint4(B[0], B[1], B[2], B[3]);
                    };
                    const unsigned char* rp = raw + wsub * kRawPitch + 8 * wk;
                    uint32_t* dst = tile + wsub * kRS + 4 * wk;
                    if (pitch15 == 0) {                                        // every row starts at the same offset inside its 16 bytes
                        rp += s0;
                        for (int r = wsub; r < trows; r += 3, rp += 3 * kRawPitch, dst += 3 * kRS) widen_row(rp, dst);
                    } else {
                        uint32_t s = (s0 + (uint32_t)wsub * pitch15) & 3u;
                        const uint32_t sstep = (3u * pitch15) & 3u;
                        for (int r = wsub; r < trows; r += 3, rp += 3 * kRawPitch, dst += 3 * kRS, s = (s + sstep) & 3u) widen_row(rp + s, dst);
                    }
                }
            }
            __syncwarp();
            // ---- the raw buffer is free: fetch the next chunk (of this cell row or of the next one)
            {
                int ny = -1, nt = 0;
                if (ya + kR < eh) { ny = y0 + ya + kR; nt = min(eh - ya - kR, kR) + 6; }
                else if (ci + 1 < ci1) {
                    const int y1 = y0 + hc, eh1 = min(y1 + hc + 6, maxBY) - y1 - 6;
                    if (eh1 > 0) { ny = y1; nt = min(eh1, kR) + 6; }
                }
                if (ny >= 0) fetch(ny, nt);
            }
            // ---- scores, NMS of the row above, survivor record; rows ya .. ya+nr-1, two per step
            uint32_t* st = stash + (ya >> 1) * 32;                     // rows (ya-2, ya-1) = slot index ya/2: indices are shifted by one
            auto score = [&](const uint32_t* a, const uint32_t* b) -> uint32_t {
                uint32_t E[16];
                E[0] = b[6 * kRS + 1];    //  ( 0, 3)
                E[1] = a[6 * kRS + 2];    //  ( 1, 3)
                E[2] = b[5 * kRS + 2];    //  ( 2, 2)
                E[3] = a[4 * kRS + 3];    //  ( 3, 1)
                E[4] = a[3 * kRS + 3];    //  ( 3, 0)
                E[5] = a[2 * kRS + 3];    //  ( 3,-1)
                E[6] = b[1 * kRS + 2];    //  ( 2,-2)
                E[7] = a[0 * kRS + 2];    //  ( 1,-3)
                E[8] = b[0 * kRS + 1];    //  ( 0,-3)
                E[9] = a[0 * kRS + 1];    //  (-1,-3)
                E[10] = b[1 * kRS + 0];   //  (-2,-2)
                E[11] = a[2 * kRS + 0];   //  (-3,-1)
                E[12] = a[3 * kRS + 0];   //  (-3, 0)
                E[13] = a[4 * kRS + 0];   //  (-3, 1)
                E[14] = b[5 * kRS + 0];   //  (-2, 2)
                E[15] = a[6 * kRS + 1];   //  (-1, 3)
                const uint32_t v = b[3 * kRS + 1];
                uint32_t M1, M2;
                fast_network(E, M1, M2);
                const uint32_t t = hmax2(hsub2(M1, v), hsub2(v, M2));     // score + 1, signed
                return hfma2_relu(t, lm, bias);                           // max(score - minTh + 1, 0); 0 in the invalid halves
            };
            // strict 3x3 maximum of row mm between rows uu and dd: mm where it survives, else 0
            auto nms = [&](const uint32_t uu, const uint32_t mm, const uint32_t dd) -> uint32_t {
                const uint32_t Uw = hmax2(uu, dd), Cw = hmax2(Uw, mm);
                const uint32_t Lw = __shfl_up_sync(0xffffffffu, Cw, 1), Rw = __shfl_down_sync(0xffffffffu, Cw, 1);
                const uint32_t nb = hmax2(hmax2(__byte_perm(Lw, Cw, selL), __byte_perm(Cw, Rw, selR)), Uw);
                const uint32_t g = hsetgt2(mm, nb);
                cntAll = hadd2(cntAll, g);
                const uint32_t s = hmul2(mm, g);
                cntA = hadd2(cntA, hsetge2(s, iniV));
                return s;
            };
            // word of a row pair: even-row survivors se, odd-row survivors so (scores are fp16 denormals: the bits are the integers)
            auto record = [&](uint32_t* slot, const uint32_t se, const uint32_t so) {
                uint32_t w;
                asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(w) : "r"(so), "r"(se));
                *slot = w;
                smask = smask * 2u + min(w, 1u);                       // newest record in bit 0: one VIMNMX + one IMAD
            };
            int r = 0;
#pragma unroll (kFastUnroll)
            for (; r + 2 <= nr; r += 2, st += 32) {
                const uint32_t* a = po + r * kRS;
                const uint32_t d0 = score(a, a + kCopyW);
                const uint32_t d1 = score(a + kRS, a + kRS + kCopyW);
                const uint32_t so = nms(u, m, d0);                        // survivors of the odd row above d0
                record(st, carry, so);
                carry = nms(m, d0, d1);                                   // survivors of the even row d0
                u = d0; m = d1;
            }
            if (r < nr) {                                                 // single last row (even index): only at the end of a cell row
                const uint32_t* a = po + r * kRS;
                const uint32_t d0 = score(a, a + kCopyW);
                const uint32_t so = nms(u, m, d0);
                record(st, carry, so);
                st += 32;
                // the row below is outside the cell: finish now, this row is the even row of its pair
                const uint32_t se = nms(m, d0, 0u);
                record(st, se, 0u);
            } else if (ya + kR >= eh) {                                   // the cell row ended on an odd row
                const uint32_t so = nms(u, m, 0u);
                record(st, carry, so);
            }
        }
        __syncwarp();

        // ---- per-cell cut-off (:857-861), one atomic per cell row and band, every lane writes its own survivors
        // survivors of iniThFAST per cell; the fp16 counts are exact (at most one per row pair and column)
        const bool aLo = (cntA & 0x7fffu) != 0, aHi = (cntA >> 16 & 0x7fffu) != 0;
        const uint32_t any0 = __ballot_sync(0xffffffffu, (aLo && clo == 0) || (aHi && chi == 0));
        const uint32_t any1 = __ballot_sync(0xffffffffu, (aLo && clo == 1) || (aHi && chi == 1));
        const bool iniLo = clo ? any1 != 0 : any0 != 0, iniHi = chi ? any1 != 0 : any0 != 0;   // cut-off of each half's cell
        const int nAllLo = __half2int_rn(__ushort_as_half((unsigned short)(cntAll & 0xffffu))), nAllHi = __half2int_rn(__ushort_as_half((unsigned short)(cntAll >> 16)));
        const int nALo = __half2int_rn(__ushort_as_half((unsigned short)(cntA & 0xffffu))), nAHi = __half2int_rn(__ushort_as_half((unsigned short)(cntA >> 16)));
        const int mine = (iniLo ? nALo : nAllLo) + (iniHi ? nAHi : nAllHi);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&io.cand_count[frame * ORB_MAX_LEVELS + l], total);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int slot = base + incl - mine;
            uint32_t* out = io.cand + (size_t)frame * plan.cand_per_frame + L.cand_off;
            uint32_t* outp = out + slot;
            const int room = L.cand_cap - slot;                         // entries this lane may still write (the capacity is a proven bound)
            int wrote = 0;
            const int xb = X0 + t0 + 3, yb = ci * hc + 3 - 2, sb = plan.minTh - 1;   // slot index i holds rows 2i-2, 2i-1
            auto emit = [&](const uint32_t v, const int hi, const int i) {
                const int odd = v > 255u ? 1 : 0;
                const int sc = (int)(odd ? v >> 8 : v);
                if (sc >= iniBias || !(hi ? iniHi : iniLo)) {
                    // border-frame coordinates (src/ORBextractor.cc:868-869): cell-local + (j*wCell, i*hCell)
                    if (wrote < room) outp[wrote] = orb_pack(xb + hi, yb + 2 * i + odd, sc + sb);
                    ++wrote;
                }
            };
            const int nrec = ((eh + 1) >> 1) + 1;                      // records written: the dummy slot 0 and one per row pair
            while (smask) {
                const int i = nrec - __ffs((int)smask);                // bit b is the record written b steps before the last
                smask &= smask - 1;
                const uint32_t w = stash[32 * i];
                const uint32_t lo = w & 0xffffu;
                const int hi = lo == 0u ? 1 : 0;
                emit(hi ? w >> 16 : lo, hi, i);
                if (lo != 0u && (w >> 16) != 0u) emit(w >> 16, 1, i);  // both halves: only the lane that straddles two cells
            }
        }
        __syncwarp();                                                   // the next cell row reuses the stash
    }
}

size_t orb_fast_smem_bytes(const OrbPlan& plan)
{
    return (size_t)ORB_FAST_WPB * fast_warp_bytes(plan.fast_stash_slots);
}

// ---- host: tensor maps.  cuTensorMapEncodeTiled comes from the driver (no link dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// u8 tensor (x, y, frame) of one level; false when the layout is not 16-byte aligned
static bool encode_level(unsigned char* out, const void* base, int w, int h, int pitch, size_t frame_stride, int frames)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || ((uintptr_t)base & 15u) || (pitch & 15) || (frame_stride & 15u) || pitch < w) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames };
    const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)(frames > 1 ? frame_stride : (size_t)pitch * h + 16 & ~(size_t)15) };
    const cuuint32_t box[3] = { (cuuint32_t)kBoxW, (cuuint32_t)kTileRows, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    alignas(64) CUtensorMap m;
    if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    static_assert(sizeof(CUtensorMap) == 128, "descriptor size");
    memcpy(out, &m, 128);
    return true;
}

static void fast_maps_update(const OrbPlan& plan, const OrbBatch& io, int batch, OrbFastMaps* M)
{
    if (M->key_img0 == io.img0 && M->key_pyr == io.pyr && M->key_stride == io.img0_stride && M->key_pitch == io.img0_pitch && M->key_batch == batch &&
        M->key_w == plan.w && M->key_h == plan.h && M->key_levels == plan.nlevels) return;
    M->use = 0;
    for (int l = 0; l < plan.nlevels; ++l) {
        const OrbLevel& L = plan.lv[l];
        const bool ok = l == 0 ? encode_level(M->map[0], io.img0, L.w, L.h, io.img0_pitch, io.img0_stride, batch)
                               : encode_level(M->map[l], io.pyr + L.img_off, L.w, L.h, L.pitch, plan.pyr_bytes, batch);
        if (ok) M->use |= 1u << l;
    }
    M->key_img0 = io.img0; M->key_pyr = io.pyr; M->key_stride = io.img0_stride; M->key_pitch = io.img0_pitch; M->key_batch = batch;
    M->key_w = plan.w; M->key_h = plan.h; M->key_levels = plan.nlevels;
}

void fast_maps_prepare(const OrbPlan& plan, const OrbBatch& io, int batch, OrbFastMaps* maps) { fast_maps_update(plan, io, batch, maps); }

cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st, OrbFastMaps* maps, int level_lo, int level_hi)
{
    if (level_hi > plan.nlevels) level_hi = plan.nlevels;
    if (plan.total_cells == 0) return cudaSuccess;
    const size_t smem = orb_fast_smem_bytes(plan);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_bands, cudaFuncAttributeMaxDynamicSharedMemorySize, ORB_SMEM_OPTIN);   // a constant: the attribute is per-function state shared by all host threads (a per-launch value races)
        if (e != cudaSuccess) return e;
    }
    // Cell rows per block: whole band columns once the batch alone fills the GPU several times over (the
    // per-band setup is then paid once per column), single cell rows for a few frames (latency).
    int maxrows = 1;
    for (int l = 0; l < plan.nlevels; ++l) if (plan.lv[l].ncy > maxrows) maxrows = plan.lv[l].ncy;
    FastGrid fg;
    fg.rows_per_block = batch >= ORB_FAST_FULLCOL ? maxrows : batch >= 16 ? ORB_FAST_MIDROWS : 1;
    int n = 0;
    for (int l = 0; l < plan.nlevels; ++l) {
        fg.first[l] = n;
        if (l < level_lo || l >= level_hi) continue;          // levels outside the range get no blocks
        const int bpr = (plan.lv[l].fbands + ORB_FAST_WPB - 1) / ORB_FAST_WPB;
        n += bpr * ((plan.lv[l].ncy + fg.rows_per_block - 1) / fg.rows_per_block);
    }
    for (int l = plan.nlevels; l <= ORB_MAX_LEVELS; ++l) fg.first[l] = n;
    if (n == 0) return cudaSuccess;
    OrbFastMaps local;                                       // callers without a cache: encoded per launch
    if (!maps) maps = &local;
    fast_maps_update(plan, io, batch, maps);
    fg.tensor_levels = maps->use;
    static_assert(sizeof(FastMapsArg) == sizeof(maps->map), "descriptor block");
    k_fast_bands<<<dim3(n, batch), kNT, smem, st>>>(plan, io, fg, *reinterpret_cast<const FastMapsArg*>(maps->map));
    return cudaGetLastError();
}
