/*
 * xtc.c — minimal GROMACS XTC reader (whole files into memory), TEST INFRASTRUCTURE ONLY.
 *
 * The reference reads trajectories through groan_rs' GroupXtcReader -> molly 0.5.0
 * (reference call site: src/analysis/common.rs:281-304; Cargo.lock:955), neither vendored.
 * This restates the published xdrfile "xdr3dfcoord" decompression so that the reference's own
 * XTC fixtures (tests/files/ua.xtc, pcpepg_selected.xtc) can drive the oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const int magicints[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
    645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768,
    41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255,
    1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085,
    16777216};
#define FIRSTIDX 9

typedef struct { const unsigned char *p; size_t cnt; unsigned int lastbits, lastbyte; } BitReader;

static int receivebits(BitReader *b, int num_of_bits) {
    int num = 0;
    int mask = (int)((1u << num_of_bits) - 1u);
    unsigned int lastbits = b->lastbits, lastbyte = b->lastbyte;
    while (num_of_bits >= 8) {
        lastbyte = (lastbyte << 8) | b->p[b->cnt++];
        num |= (int)((lastbyte >> lastbits) << (num_of_bits - 8));
        num_of_bits -= 8;
    }
    if (num_of_bits > 0) {
        if (lastbits < (unsigned int)num_of_bits) {
            lastbits += 8;
            lastbyte = (lastbyte << 8) | b->p[b->cnt++];
        }
        lastbits -= (unsigned int)num_of_bits;
        num |= (int)((lastbyte >> lastbits) & ((1u << num_of_bits) - 1u));
    }
    num &= mask;
    b->lastbits = lastbits;
    b->lastbyte = lastbyte;
    return num;
}

static void receiveints(BitReader *b, int num_of_ints, int num_of_bits, const unsigned int sizes[], int nums[]) {
    int bytes[32];
    int i, j, num_of_bytes = 0, p, num;
    bytes[1] = bytes[2] = bytes[3] = 0;
    while (num_of_bits > 8) { bytes[num_of_bytes++] = receivebits(b, 8); num_of_bits -= 8; }
    if (num_of_bits > 0) bytes[num_of_bytes++] = receivebits(b, num_of_bits);
    for (i = num_of_ints - 1; i > 0; i--) {
        num = 0;
        for (j = num_of_bytes - 1; j >= 0; j--) {
            num = (num << 8) | bytes[j];
            p = (int)((unsigned int)num / sizes[i]);
            bytes[j] = p;
            num = num - p * (int)sizes[i];
        }
        nums[i] = num;
    }
    nums[0] = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
}

static int sizeofint(unsigned int size) {
    unsigned int num = 1;
    int bits = 0;
    while (size >= num && bits < 32) { bits++; num <<= 1; }
    return bits;
}

static int sizeofints(int num_of_ints, const unsigned int sizes[]) {
    unsigned int num_of_bytes = 1, num_of_bits = 0, bytes[32], bytecnt, tmp, num;
    bytes[0] = 1;
    for (int i = 0; i < num_of_ints; i++) {
        tmp = 0;
        for (bytecnt = 0; bytecnt < num_of_bytes; bytecnt++) {
            tmp = bytes[bytecnt] * sizes[i] + tmp;
            bytes[bytecnt] = tmp & 0xff;
            tmp >>= 8;
        }
        while (tmp != 0) { bytes[bytecnt++] = tmp & 0xff; tmp >>= 8; }
        num_of_bytes = bytecnt;
    }
    num = 1;
    num_of_bytes--;
    while (bytes[num_of_bytes] >= num) { num_of_bits++; num *= 2; }
    return (int)(num_of_bits + num_of_bytes * 8);
}

static int32_t rd_i32(const unsigned char *p) { return (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]); }
static float rd_f32(const unsigned char *p) { uint32_t u = (uint32_t)rd_i32(p); float f; memcpy(&f, &u, 4); return f; }

/* Parse one frame at data[pos..]; returns bytes consumed or 0 on EOF / -1 on error.
 * xyz may be NULL (scan only). box9: 3x3 matrix. */
static long parse_frame(const unsigned char *data, size_t size, size_t pos, int *natoms, int *step, float *time,
                        float *box9, float *prec_out, float *xyz) {
    size_t p = pos;
    if (p + 16 > size) return 0;
    int magic = rd_i32(data + p);
    if (magic != 1995 && magic != 2023) return -1;
    *natoms = rd_i32(data + p + 4);
    *step = rd_i32(data + p + 8);
    *time = rd_f32(data + p + 12);
    p += 16;
    if (p + 36 + 4 > size) return -1;
    for (int i = 0; i < 9; i++) box9[i] = rd_f32(data + p + 4 * i);
    p += 36;
    int lsize = rd_i32(data + p);
    p += 4;
    if (lsize != *natoms) return -1;
    if (lsize <= 9) {
        if (p + 12 * (size_t)lsize > size) return -1;
        if (xyz) for (int i = 0; i < 3 * lsize; i++) xyz[i] = rd_f32(data + p + 4 * i);
        *prec_out = 0.0f;
        return (long)(p + 12 * (size_t)lsize - pos);
    }
    if (p + 4 + 24 + 4 + 4 > size) return -1;
    float precision = rd_f32(data + p); p += 4;
    *prec_out = precision;
    int minint[3], maxint[3];
    for (int i = 0; i < 3; i++) minint[i] = rd_i32(data + p + 4 * i);
    p += 12;
    for (int i = 0; i < 3; i++) maxint[i] = rd_i32(data + p + 4 * i);
    p += 12;
    int smallidx = rd_i32(data + p); p += 4;
    size_t nbytes;
    if (magic == 2023) { /* 64-bit byte count */
        if (p + 8 > size) return -1;
        nbytes = ((size_t)(uint32_t)rd_i32(data + p) << 32) | (uint32_t)rd_i32(data + p + 4);
        p += 8;
    } else { nbytes = (size_t)(uint32_t)rd_i32(data + p); p += 4; }
    size_t padded = (nbytes + 3) & ~(size_t)3;
    if (p + padded > size) return -1;
    if (xyz) {
        unsigned int sizeint[3], sizesmall[3], bitsizeint[3] = {0, 0, 0};
        int bitsize;
        for (int i = 0; i < 3; i++) sizeint[i] = (unsigned int)(maxint[i] - minint[i] + 1);
        if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff) {
            for (int i = 0; i < 3; i++) bitsizeint[i] = (unsigned int)sizeofint(sizeint[i]);
            bitsize = 0;
        } else bitsize = sizeofints(3, sizeint);
        int tmp = smallidx - 1;
        tmp = (FIRSTIDX > tmp) ? FIRSTIDX : tmp;
        int smaller = magicints[tmp] / 2;
        int smallnum = magicints[smallidx] / 2;
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned int)magicints[smallidx];
        /* the bit reader may look one byte ahead: copy into a padded scratch buffer */
        unsigned char *scratch = (unsigned char *)calloc(padded + 8, 1);
        memcpy(scratch, data + p, nbytes);
        BitReader br = {scratch, 0, 0, 0};
        float inv_precision = 1.0f / precision;
        int run = 0, i = 0;
        float *lfp = xyz;
        int thiscoord[3], prevcoord[3];
        while (i < lsize) {
            if (bitsize == 0) {
                thiscoord[0] = receivebits(&br, (int)bitsizeint[0]);
                thiscoord[1] = receivebits(&br, (int)bitsizeint[1]);
                thiscoord[2] = receivebits(&br, (int)bitsizeint[2]);
            } else receiveints(&br, 3, bitsize, sizeint, thiscoord);
            i++;
            thiscoord[0] += minint[0]; thiscoord[1] += minint[1]; thiscoord[2] += minint[2];
            prevcoord[0] = thiscoord[0]; prevcoord[1] = thiscoord[1]; prevcoord[2] = thiscoord[2];
            int flag = receivebits(&br, 1);
            int is_smaller = 0;
            if (flag == 1) {
                run = receivebits(&br, 5);
                is_smaller = run % 3;
                run -= is_smaller;
                is_smaller--;
            }
            if (run > 0) {
                for (int k = 0; k < run; k += 3) {
                    receiveints(&br, 3, smallidx, sizesmall, thiscoord);
                    i++;
                    thiscoord[0] += prevcoord[0] - smallnum;
                    thiscoord[1] += prevcoord[1] - smallnum;
                    thiscoord[2] += prevcoord[2] - smallnum;
                    if (k == 0) {
                        /* first and second atom are interchanged (water compression trick) */
                        for (int c = 0; c < 3; c++) { int t = thiscoord[c]; thiscoord[c] = prevcoord[c]; prevcoord[c] = t; }
                        *lfp++ = prevcoord[0] * inv_precision;
                        *lfp++ = prevcoord[1] * inv_precision;
                        *lfp++ = prevcoord[2] * inv_precision;
                    } else {
                        prevcoord[0] = thiscoord[0]; prevcoord[1] = thiscoord[1]; prevcoord[2] = thiscoord[2];
                    }
                    *lfp++ = thiscoord[0] * inv_precision;
                    *lfp++ = thiscoord[1] * inv_precision;
                    *lfp++ = thiscoord[2] * inv_precision;
                }
            } else {
                *lfp++ = thiscoord[0] * inv_precision;
                *lfp++ = thiscoord[1] * inv_precision;
                *lfp++ = thiscoord[2] * inv_precision;
            }
            smallidx += is_smaller;
            if (is_smaller < 0) {
                smallnum = smaller;
                smaller = (smallidx > FIRSTIDX) ? magicints[smallidx - 1] / 2 : 0;
            } else if (is_smaller > 0) {
                smaller = smallnum;
                smallnum = magicints[smallidx] / 2;
            }
            sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned int)magicints[smallidx];
            if (lfp - xyz > 3 * (long)lsize) { free(scratch); return -1; }
        }
        free(scratch);
    }
    return (long)(p + padded - pos);
}

static unsigned char *slurp(const char *path, size_t *size) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *d = (unsigned char *)malloc((size_t)n + 1);
    if (fread(d, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(d); return NULL; }
    fclose(f);
    *size = (size_t)n;
    return d;
}

/* count frames and atoms */
int xtc_scan(const char *path, int *natoms, int *nframes) {
    size_t size;
    unsigned char *d = slurp(path, &size);
    if (!d) return -1;
    size_t pos = 0;
    int n = 0, na = 0, step;
    float time, box9[9], prec;
    for (;;) {
        long used = parse_frame(d, size, pos, &na, &step, &time, box9, &prec, NULL);
        if (used == 0) break;
        if (used < 0) { free(d); return -2; }
        pos += (size_t)used;
        n++;
    }
    free(d);
    *natoms = na;
    *nframes = n;
    return 0;
}

/* read every frame: xyz[nframes][natoms][3], box9[nframes][9], time[nframes], step[nframes] */
int xtc_read(const char *path, int natoms, int nframes, float *xyz, float *box9, float *time, int *step, float *prec) {
    size_t size;
    unsigned char *d = slurp(path, &size);
    if (!d) return -1;
    size_t pos = 0;
    for (int f = 0; f < nframes; f++) {
        int na;
        long used = parse_frame(d, size, pos, &na, &step[f], &time[f], box9 + 9 * f, prec, xyz + (size_t)f * natoms * 3);
        if (used <= 0 || na != natoms) { free(d); return -2; }
        pos += (size_t)used;
    }
    free(d);
    return 0;
}
