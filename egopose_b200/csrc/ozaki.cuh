// Internal interface of the int8-tensor-core float64 GEMM (ozaki.cu), shared with the chunked MLP driver (oz_mlp.cu).
#pragma once
#include "common.cuh"

namespace egp {
namespace oz {

constexpr int BM = 128;             // output rows per tile (tcgen05 M)
#ifndef OZ_BK
#define OZ_BK 64
#endif
// Contraction bytes per stage row (32: one tcgen05.mma.kind::i8 K step, 32-byte swizzle; 64: two, 64-byte swizzle).  The TMA
// unit serves about one box ROW per clock whatever its width, and a 128 x 80 x K tile asks for (128 + 80) S K / BK operand rows:
// at BK = 32 that is 12.5 k requests per 17 k-clock tile (K = 320) and the producer's AND the epilogue's store requests queue
// behind each other (the store instruction itself stalled ~700 clocks); 64-byte rows halve the request count (two stages of
// 80 KB instead of four of 40 KB: same bytes in flight).
constexpr int BK = OZ_BK;
#ifndef OZ_RADIX_BITS
#define OZ_RADIX_BITS 7
#endif
// Digit width of the slices.  7 (default): signed round-to-nearest digits in [-64, 64], all slices signed.
// 8 (experiment, -DOZ_RADIX_BITS=8): the two's-complement bytes of the fixed-point value (top slice signed, the others
// unsigned; tcgen05 kind::i8 takes the signedness per operand and instruction).  Measured on B200: bit-exact as well,
// but the one-sided truncation of unsigned digits accumulates linearly over the contraction, so at equal slice count it
// is ~200x LESS accurate than the centred radix-128 digits despite the extra bit per slice (err / max|C| 6.8e-10 at S = 5
// vs 3.1e-10 for radix 128 at S = 5 and 3.8e-12 at S = 6) - no gain, kept only as a documented negative result.
constexpr int RB = OZ_RADIX_BITS;
constexpr int MAX_S = RB == 8 ? 6 : (BK == 64 ? 7 : 8);     // 8 slices x two 64-byte stages do not fit shared memory (7 = DGEMM rounding level)
constexpr int EOFF = 2 - 2 * RB;    // C = 2^(ea + eb + EOFF) * sum_d acc_d 2^(-RB d)
// k blocks of one int32 accumulation: S pairs x K x max |digit product| < 2^31
inline long long max_kblocks(int S) { return RB == 8 ? (((1LL << 31) / (65025LL * S * BK)) & ~1LL) : ((16384 / S * 32 / BK) & ~1LL); }

inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// f64 [m][k] (ld ldx) -> int8 [S][m][kp] + exps[m]; optional column abs-max (bit patterns, atomicMax) of x
int slice_rows(const double *x, long long m, int k, long long ldx, int S, int8_t *out, int kp, int32_t *exps,
               unsigned long long *colmax, cudaStream_t st);
int col_absmax(const double *x, long long n, int f, long long ldx, unsigned long long *colmax, cudaStream_t st);
// f64 [n][f] -> int8 [S][f (+1 row of ones)][np], exps[f (+1)] from colmax
int slice_colsT(const double *x, long long n, int f, long long ldx, int S, const unsigned long long *colmax, int8_t *out,
                long long np, int32_t *exps, int ones_row, cudaStream_t st);

// both orientations from one read of x, given the row abs-max high words (rowmax [n]) and the column abs-max bit patterns:
// row slices int8 [S][n][kp] + expsR[n] (kp multiple of 32), transposed slices int8 [S][f (+1)][np] + expsT[f (+1)]
int slice_both(const double *x, long long n, int f, long long ldx, int S, const uint32_t *rowmax, const unsigned long long *colmax,
               int8_t *outR, int kp, int32_t *expsR, int8_t *outT, long long np, int32_t *expsT, int ones_row, cudaStream_t st);

struct GemmOut {
    double *C = nullptr;            // [m][ldc]
    long long ldc = 0;
    const double *bias = nullptr;   // [n] added before relu
    int relu = 0;
    const double *mask = nullptr;   // [m][ldm]: C = mask > 0 ? C : 0 (relu backward)
    long long ldm = 0;
    // split-K (weight gradients): raw partial sums [splits][m][ldp] go to work; the caller reduces them
    double *work = nullptr;
    long long work_bytes = 0;
    int force_splits = 0;           // > 0: use exactly this many splits (and write partials even when 1)
    // abs-maxima of the final output for its slicers (optional; zeroed by the caller; not on the split-K path):
    uint32_t *rowmax = nullptr;            // [m] high words (atomicMax)
    unsigned long long *colmax = nullptr;  // [n] bit patterns, high word << 32 (atomicMax) - same encoding as slice_rows' colmax
    // relu sign bits, one 64-bit word per (row, half n-tile): written by a relu GEMM (relu_bits), consumed INSTEAD of the
    // float64 mask by the relu-backward GEMM over the same [m][n] shape (mask_bits): 8 bytes per thread and tile in place of
    // its columns as doubles.  Both [m][gemm_bits_words(n, S)] words (one per epilogue warp of the row).
    unsigned long long *relu_bits = nullptr;
    const unsigned long long *mask_bits = nullptr;
    int splits_used = 0;            // out
    long long ldp = 0;              // out
};
long long gemm_work_bytes(long long m, int n, long long kp, int force_splits);
int choose_splits(long long m, int n, long long kp, int S);
long long gemm_tiles(long long m, int n, int S);      // output tiles of one split
int gemm_ntiles(int n, int S);                        // n-tiles (columns of tiles) of an [m][n] output
int gemm_bits_words(int n, int S);                    // 64-bit words per row of the relu sign bits of an [m][n] output
// C = A B^T from slices A [S][m][kp], B [S][n][kp]
int gemm(const int8_t *a, const int32_t *ea, long long m, const int8_t *b, const int32_t *eb, int n, long long kp, int S,
         GemmOut &o, cudaStream_t st);

}  // namespace oz
}  // namespace egp
