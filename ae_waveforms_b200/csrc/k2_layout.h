// k2_layout.h -- tile geometry and coefficient-table layout shared by the K2
// kernel (k2_filtfilt.cu) and the host-side table builder (k2_tables.cpp).
#pragma once

#define K2_C 13                      // samples per thread chunk (odd: conflict-free stride in shared memory)
#ifndef K2_NT
#define K2_NT 128                    // threads per block (-DK2_NT=256 rebuilds round 1's geometry: 3,328-sample tiles,
#endif                               //   two blocks per SM)
#define K2_NW (K2_NT / 32)
#define K2_T (K2_C * K2_NT)          // samples per tile = 1664 (13 KB)
#define K2_RUN (106496 / K2_T)        // tiles of a segment one block works through in a row (k2_pass): 64 resp. 32
#define K2_BPS (512 / K2_NT)         // resident blocks per SM: 512 threads at 128 registers fill the register file
#define K2_STAGES 4                  // ring of tile buffers per block: loading / recursion / correction / storing
#define K2_MAXPAD 15                 // filtfilt padlen for order 4

// Table of doubles (device copy is read by every block once per launch).
// The tiled pass runs the filter in MODAL form: up to two complex first-order
// sections u' = p u + x (a conjugate pole pair is carried once with weight 2),
// y = d x + Re(g1 u1) + Re(g2 u2).  State vector s = (u1.re, u1.im, u2.re, u2.im).
#define K2_TAB_B 0                   // b[0..4]   (exact mode)
#define K2_TAB_A 5                   // a[0..4]   (exact mode)
#define K2_TAB_ZI 10                 // zi[0..3]  scipy.signal.lfilter_zi
#define K2_TAB_ORDER 14
#define K2_TAB_PAD 15
#define K2_TAB_LOOKBACK 16           // tiles in front of a run whose samples still matter: the run's warm-up
#define K2_TAB_D 17                  // direct term b[0]
#define K2_TAB_P 20                  // p1.re p1.im p2.re p2.im
#define K2_TAB_GC 24                 // y += GC . s          (g.re, -g.im per section)
#define K2_TAB_PZ 28                 // [4][4] rows: s = PZ . z   (DF2T state -> modal state)
#define K2_TAB_PWARP 44              // p^(32 C)   (re, im) x 2
#define K2_TAB_PTILE 48              // p^T
#define K2_TAB_PK 52                 // [5][4]  p^(C 2^k)
#define K2_TAB_PLANE 72              // [32][4] p^(C l)
#define K2_TAB_G (K2_TAB_PLANE + 128)          // [C][4]  zero-input response: y_j += G[j] . s
#define K2_TAB_PWP (K2_TAB_G + 4 * K2_C)        // [NW+1][4] p^(32 C m), m = 0 .. NW
#define K2_TAB_LEN (K2_TAB_PWP + 4 * (K2_NW + 1))
