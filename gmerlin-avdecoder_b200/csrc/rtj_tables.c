/*
 * rtj_tables.c -- host-side dequantisation tables of the RTjpeg decoder.
 *
 * Replaces, on the host, RTjpeg_calc_tbls (lib/RTjpeg.c:2344-2369),
 * RTjpeg_set_tables (:2380-2395) and RTjpeg_idct_init (:1208-1217) of the
 * reference.  Tables are a pure function of the quality byte, so the batch
 * context builds all 255 of them once and keeps them in device memory.
 */
#include "rtj_common.h"

#include <string.h>

/* zig-zag position -> raster index; the reference's order is the transpose of
 * JPEG's (position 1 is row 1 / column 0), lib/RTjpeg.c:59-74. */
const uint8_t rtj_zigzag[64] = {
     0,  8,  1,  2,  9, 16, 24, 17, 10,  3,  4, 11, 18, 25, 32, 40,
    33, 26, 19, 12,  5,  6, 13, 20, 27, 34, 41, 48, 56, 49, 42, 35,
    28, 21, 14,  7, 15, 22, 29, 36, 43, 50, 57, 58, 51, 44, 37, 30,
    23, 31, 38, 45, 52, 59, 60, 53, 46, 39, 47, 54, 61, 62, 55, 63
};

/* ITU-T T.81 Annex K quantisation tables (lib/RTjpeg.c:87-107), raster order. */
static const uint8_t annexk_luma[64] = {
    16, 11, 10, 16,  24,  40,  51,  61,
    12, 12, 14, 19,  26,  58,  60,  55,
    14, 13, 16, 24,  40,  57,  69,  56,
    14, 17, 22, 29,  51,  87,  80,  62,
    18, 22, 37, 56,  68, 109, 103,  77,
    24, 35, 55, 64,  81, 104, 113,  92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103,  99
};
static const uint8_t annexk_chroma[64] = {
    17, 18, 24, 47, 99, 99, 99, 99,
    18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99
};

/* AAN post-scale factors, Q32 fixed point (lib/RTjpeg.c:76-85).  The 8x8 table
 * is symmetric and its rows 0 and 4 coincide; only the upper triangle is
 * stored, row r holding columns r..7. */
static const uint64_t aan_upper[36] = {
    0x100000000ULL, 0x163140200ULL, 0x14e7c0600ULL, 0x12d07fa00ULL, 0x100000000ULL, 0x0c9240700ULL, 0x08a8c0500ULL, 0x046a00180ULL,
    0x1ec83fe00ULL, 0x1cff00200ULL, 0x1a187f800ULL, 0x163140200ULL, 0x116fc0400ULL, 0x0c02bfa00ULL, 0x061f7f900ULL,
    0x1b503fc00ULL, 0x189500000ULL, 0x14e7c0600ULL, 0x106cbfc00ULL, 0x0b503fb00ULL, 0x05c480600ULL,
    0x161f7f800ULL, 0x12d07fa00ULL, 0x0ec83fd00ULL, 0x0a2e80800ULL, 0x0530c0280ULL,
    0x100000000ULL, 0x0c9240700ULL, 0x08a8c0500ULL, 0x046a00180ULL,
    0x09e080700ULL, 0x06cdc0100ULL, 0x037800200ULL,
    0x04afc0500ULL, 0x026380040ULL,
    0x0137c02a0ULL
};

static uint64_t aan_factor(int r, int c)
{
    if (r > c) { int t = r; r = c; c = t; }
    /* rows before r hold 8, 7, ... entries */
    int base = r * 8 - (r * (r - 1)) / 2;
    return aan_upper[base + (c - r)];
}

/* How many AC coefficients after DC travel as plain signed bytes: the leading
 * zig-zag positions whose unscaled multiplier is <= 8 (lib/RTjpeg.c:2362-2367).
 * The reference scans without an upper bound; a table that never exceeds 8
 * yields 63 here. */
static int raw_prefix(const int32_t *unscaled)
{
    int n = 0;
    while (n < 63 && unscaled[rtj_zigzag[n + 1]] <= 8) n++;
    return n;
}

static void scale_by_aan(int32_t *t)
{
    for (int i = 0; i < 64; i++)
        t[i] = (int32_t)(((uint64_t)(uint32_t)t[i] * aan_factor(i >> 3, i & 7)) >> 32);
}

void rtj_table_from_quality(int Q, rtj_host_table *out)
{
    if (Q < 1) Q = 1;
    if (Q > 255) Q = 255;
    const uint64_t scaled_q = (uint64_t)Q << 25;    /* 32-bit fixed point: 255 -> ~2.0 */
    for (int i = 0; i < 64; i++) {
        int32_t ql = (int32_t)((scaled_q / ((uint64_t)annexk_luma[i] << 16)) >> 3);
        int32_t qc = (int32_t)((scaled_q / ((uint64_t)annexk_chroma[i] << 16)) >> 3);
        out->liqt[i] = 65536 / ((ql ? ql : 1) << 3);
        out->ciqt[i] = 65536 / ((qc ? qc : 1) << 3);
    }
    out->lb8 = raw_prefix(out->liqt);
    out->cb8 = raw_prefix(out->ciqt);
    scale_by_aan(out->liqt);
    scale_by_aan(out->ciqt);
}

/* The encoder's quantiser multipliers for quality Q, raster order, luma then chroma, plus the raw-prefix lengths:
 * RTjpeg_calc_tbls (lib/RTjpeg.c:2344-2369) followed by RTjpeg_dct_init (:277-286). */
void rtj_encoder_table_from_quality(int Q, int32_t qt[128], int *lb8, int *cb8)
{
    if (Q < 1) Q = 1;
    if (Q > 255) Q = 255;
    const uint64_t scaled_q = (uint64_t)Q << 25;
    int32_t lpre[64], cpre[64];
    for (int i = 0; i < 64; i++) {
        int32_t ql = (int32_t)((scaled_q / ((uint64_t)annexk_luma[i] << 16)) >> 3);
        int32_t qc = (int32_t)((scaled_q / ((uint64_t)annexk_chroma[i] << 16)) >> 3);
        lpre[i] = 65536 / ((ql ? ql : 1) << 3);
        cpre[i] = 65536 / ((qc ? qc : 1) << 3);
        ql = (65536 / lpre[i]) >> 3;
        qc = (65536 / cpre[i]) >> 3;
        qt[i] = (int32_t)(((uint64_t)ql << 32) / aan_factor(i >> 3, i & 7));
        qt[64 + i] = (int32_t)(((uint64_t)qc << 32) / aan_factor(i >> 3, i & 7));
    }
    *lb8 = raw_prefix(lpre);
    *cb8 = raw_prefix(cpre);
}

void rtjgpu_raw_tables_for_quality(int Q, uint32_t raw[128])
{
    if (Q < 1) Q = 1;
    if (Q > 255) Q = 255;
    const uint64_t scaled_q = (uint64_t)Q << 25;
    for (int i = 0; i < 64; i++) {
        int32_t ql = (int32_t)((scaled_q / ((uint64_t)annexk_luma[i] << 16)) >> 3);
        int32_t qc = (int32_t)((scaled_q / ((uint64_t)annexk_chroma[i] << 16)) >> 3);
        raw[i] = (uint32_t)(65536 / ((ql ? ql : 1) << 3));
        raw[64 + i] = (uint32_t)(65536 / ((qc ? qc : 1) << 3));
    }
}

void rtj_table_from_raw(const uint32_t raw[128], rtj_host_table *out)
{
    for (int i = 0; i < 64; i++) {
        out->liqt[i] = (int32_t)raw[i];
        out->ciqt[i] = (int32_t)raw[64 + i];
    }
    out->lb8 = raw_prefix(out->liqt);
    out->cb8 = raw_prefix(out->ciqt);
    scale_by_aan(out->liqt);
    scale_by_aan(out->ciqt);
}

void rtj_table_to_device_layout(const rtj_host_table *in, rtj_dev_table *out)
{
    memset(out, 0, sizeof(*out));
    for (int k = 0; k < 64; k++) {
        out->iq[0][k] = in->liqt[rtj_zigzag[k]];
        out->iq[1][k] = in->ciqt[rtj_zigzag[k]];
    }
    out->bt8[0] = in->lb8;
    out->bt8[1] = in->cb8;
}
