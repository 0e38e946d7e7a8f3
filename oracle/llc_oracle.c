/*
 * llc_oracle.c -- TEST INFRASTRUCTURE ONLY (see llc_oracle.h for the rules).
 *
 * Plain-C restatement of the reference's LZ4 / Snappy RAP path.  Written from the
 * behavioural description in SURVEY.md Appendix A and checked against the compiled
 * reference; it is deliberately simple (byte loops, no SIMD, no wild copies) because
 * its only job is to say what the right bytes are.
 * Citations are file:line under /root/reference.
 */
#include "llc_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- little helpers */
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void st16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static inline void st32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static inline void st64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

/* ---------------------------------------------------------------- RAP arithmetic */
/* threads/threads.c:55-88 */
int orc_partition_count(int64_t n, int window, int factor, int max_threads)
{
    int64_t chunk = (int64_t)window * factor;
    if (n < chunk) return 1;
    int64_t parts = n / chunk;
    int64_t rest = n % chunk;
    int64_t half = factor > 1 ? (chunk >> 1) : (window >> 1);
    if (rest >= half) parts++;
    if (max_threads > 0 && parts > max_threads) parts = max_threads;
    return (int)parts;
}

/* threads/threads.c:105-110 -- magic | frame_len | T */
static int64_t rap_write_header(uint8_t *dst, int T)
{
    int64_t frame = 16 + 12 * (int64_t)T;
    st64(dst, ORC_RAP_MAGIC);
    st32(dst + 8, (uint32_t)frame);
    st32(dst + 12, (uint32_t)T);
    return frame;
}
static void rap_write_entry(uint8_t *dst, int i, uint32_t off, uint32_t clen, uint32_t dlen)
{
    uint8_t *e = dst + 16 + 12 * (int64_t)i;
    st32(e, off); st32(e + 4, clen); st32(e + 8, dlen);
}
/* threads/threads.c:194-230: returns T (>=1) and *frame, or -1 (T field == 0). */
static int rap_read_header(const uint8_t *src, int64_t n, int64_t *frame)
{
    *frame = 0;
    if (n < 8 || ld64(src) != ORC_RAP_MAGIC) return 1;
    uint32_t flen = ld32(src + 8), T = ld32(src + 12);
    if (T == 0) return -1;
    *frame = flen;
    return (int)T;
}

int64_t orc_lz4_bound(int64_t n) { return n > 0x7E000000 ? 0 : n + n / 255 + 16; }
int64_t orc_snappy_bound(int64_t n) { return 32 + n + n / 6; }

/* ---------------------------------------------------------------- LZ4 encoder */
/* algos/lz4/lz4.c:759-777 (hash4 / hash5, little endian) */
static inline uint32_t lz4_hash(const uint8_t *p, int wide)
{
    if (wide) return (uint32_t)(((ld64(p) << 24) * 889523592379ULL) >> 52);
    return (ld32(p) * 2654435761U) >> 19;
}
/* algos/lz4/lz4.c:656-679 */
static int64_t common_prefix(const uint8_t *a, const uint8_t *b, const uint8_t *a_end)
{
    const uint8_t *s = a;
    while (a < a_end && *a == *b) { a++; b++; }
    return a - s;
}
static uint8_t *put_run_length(uint8_t *op, int64_t v) /* v already minus 15 */
{
    while (v >= 255) { *op++ = 255; v -= 255; }
    *op++ = (uint8_t)v;
    return op;
}

int64_t orc_lz4_encode_partition(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap,
                                 int emit_tail, int64_t *tail_len)
{
    const int wide = n >= 65547;               /* LZ4_64Klimit, lz4.c:694 ; table pick :2557-2563 */
    const int limited = cap >= 0;
    uint32_t *tab = (uint32_t *)calloc(8192, sizeof(uint32_t)); /* zeroed: slot 0 == position 0 */
    uint8_t *op = dst, *const olimit = dst + (limited ? cap : 0);
    int64_t anchor = 0;

    if (n == 0) { free(tab); dst[0] = 0; if (tail_len) *tail_len = 0; return 1; } /* lz4.c:2418-2428 */
    if (n >= 13) {                              /* LZ4_minLength, lz4.c:1926 */
        const int64_t mfl1 = n - 11;            /* mflimitPlusOne, lz4.c:1887 */
        const uint8_t *const mlimit = src + n - 5; /* matchlimit, lz4.c:1888 */
        int64_t ip = 1;
        tab[lz4_hash(src, wide)] = 0;           /* lz4.c:1929 */
        for (;;) {
            int64_t fwd = ip, step = 1, nb = 64, m;   /* lz4.c:1971-1977 */
            for (;;) {                          /* search, lz4.c:1979-2090 */
                int64_t cur = fwd;
                fwd += step;
                step = nb++ >> 6;
                if (fwd > mfl1) goto tail;      /* lz4.c:2001 */
                uint32_t h = lz4_hash(src + cur, wide);
                m = tab[h];
                tab[h] = (uint32_t)cur;
                if (ld32(src + m) == ld32(src + cur) && (!wide || cur - m <= 65535)) { ip = cur; break; }
            }
            while (ip > anchor && m > 0 && src[ip - 1] == src[m - 1]) { ip--; m--; } /* lz4.c:2098 */
            int64_t ll = ip - anchor;
            uint8_t *tok = op++;
            if (limited && op + ll + 8 + ll / 255 > olimit) { free(tab); return 0; }   /* lz4.c:2104-2107 */
            if (ll >= 15) { *tok = 0xF0; op = put_run_length(op, ll - 15); } else *tok = (uint8_t)(ll << 4);
            memcpy(op, src + anchor, (size_t)ll); op += ll;
            for (;;) {                          /* _next_match, lz4.c:2126-2288 */
                st16(op, (uint32_t)(ip - m)); op += 2;
                int64_t mc = common_prefix(src + ip + 4, src + m + 4, mlimit);
                ip += mc + 4;
                if (limited && op + 6 + (mc + 240) / 255 > olimit) { free(tab); return 0; } /* lz4.c:2177-2204 */
                if (mc >= 15) { *tok += 15; op = put_run_length(op, mc - 15); } else *tok += (uint8_t)mc;
                anchor = ip;
                if (ip >= mfl1) goto tail;      /* lz4.c:2227 */
                tab[lz4_hash(src + ip - 2, wide)] = (uint32_t)(ip - 2);   /* lz4.c:2230 */
                uint32_t h = lz4_hash(src + ip, wide);
                m = tab[h];
                tab[h] = (uint32_t)ip;
                if ((!wide || m + 65535 >= ip) && ld32(src + m) == ld32(src + ip)) { /* lz4.c:2278-2286 */
                    tok = op++; *tok = 0;
                    continue;
                }
                break;
            }
            ip++;                               /* lz4.c:2291 */
        }
    }
tail:
    free(tab);
    {
        int64_t run = n - anchor;
        if (!emit_tail) { *tail_len = run; return op - dst; }   /* lz4.c:2333-2338 */
        if (tail_len) *tail_len = 0;
        if (limited && op + run + 1 + (run + 255 - 15) / 255 > olimit) return 0;   /* lz4.c:2299-2311 */
        if (run >= 15) { *op++ = 0xF0; op = put_run_length(op, run - 15); } else *op++ = (uint8_t)(run << 4);
        memcpy(op, src + anchor, (size_t)run); op += run;
    }
    return op - dst;
}

int64_t orc_lz4_compress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap, int max_threads)
{
    if ((src == NULL && n != 0) || dst == NULL) return 0;
    int T = orc_partition_count(n, ORC_LZ4_WINDOW, ORC_WINDOW_FACTOR, max_threads);
    if (T == 1) {                                /* lz4.c:2674-2677, 2485-2541 */
        if (n > 0x7E000000) return 0;
        return orc_lz4_encode_partition(src, n, dst, cap >= orc_lz4_bound(n) ? -1 : cap, 1, NULL);
    }
    int64_t common = n / T, left = n % T;
    int64_t frame = rap_write_header(dst, T);
    uint8_t *out = dst + frame;
    uint8_t *body = (uint8_t *)malloc((size_t)(common + left + (common + left) / 255 + 32));
    int64_t carry = 0;            /* tail literals handed to the next partition */
    const uint8_t *carry_src = src;
    uint32_t off = (uint32_t)frame;
    for (int i = 0; i < T; i++) {                /* lz4.c:2736-2905 */
        const uint8_t *ps = src + common * i;
        int64_t pn = common + (i == T - 1 ? left : 0), tail = 0;
        int last = (i == T - 1);
        int64_t bn = orc_lz4_encode_partition(ps, pn, body, -1, last, &tail);
        if (i == 0) {
            memcpy(out, body, (size_t)bn); out += bn;
            rap_write_entry(dst, 0, off, (uint32_t)bn, (uint32_t)(pn - tail));
            off += (uint32_t)bn;
            carry = tail; carry_src = ps + pn - tail;
            continue;
        }
        if (bn == 0 && tail) {                   /* all-literal partition, lz4.c:2808-2822 */
            rap_write_entry(dst, i, off, 0, 0);
            carry += tail;                       /* carry_src unchanged: tails are contiguous */
            continue;
        }
        uint8_t *start = out;
        const uint8_t *bp = body;
        uint32_t t = *bp++;
        int64_t ll = t >> 4, nl = ll + carry;
        if (nl >= 15) {                          /* lz4.c:2825-2862 */
            int64_t acc = nl - 15;
            *out++ = (uint8_t)(0xF0 | (t & 15));
            while (acc >= 255) { *out++ = 255; acc -= 255; }
            if (ll >= 15) {
                while (*bp == 255) { *out++ = 255; bp++; }
                acc += *bp++;
                if (acc >= 255) { *out++ = 255; acc -= 255; }
            }
            *out++ = (uint8_t)acc;
        } else {
            *out++ = (uint8_t)((nl << 4) | (t & 15));
        }
        memcpy(out, carry_src, (size_t)carry); out += carry;
        int64_t rest = bn - (bp - body);
        memcpy(out, bp, (size_t)rest); out += rest;
        rap_write_entry(dst, i, off, (uint32_t)(out - start), (uint32_t)(pn - tail + carry));
        off += (uint32_t)(out - start);
        carry = tail; carry_src = ps + pn - tail;
    }
    free(body);
    (void)cap;                                   /* the reference never checks it here */
    return out - dst;
}

/* ---------------------------------------------------------------- LZ4 decoder */
/* algos/lz4/lz4.c:3806-4305, stripped of the wild-copy fast paths: same accept/reject
 * decisions for well-formed streams; hostile streams are rejected (never over-read or
 * over-written) but the negative value is not the reference's position code. */
int64_t orc_lz4_decode_partition(const uint8_t *src, int64_t clen, uint8_t *dst, int64_t cap,
                                 int is_last)
{
    int64_t ip = 0, op = 0;
    if (src == NULL || clen <= 0) return -1;
    if (cap == 0) return (clen == 1 && src[0] == 0) ? 0 : -1;     /* lz4.c:3854-3858 */
    for (;;) {
        if (ip >= clen) return -1;
        uint32_t tok = src[ip++];
        int64_t len = tok >> 4;
        if (len == 15) {
            uint32_t b;
            do { if (ip >= clen) return -1; b = src[ip++]; len += b; } while (b == 255);
        }
        if (len > clen - ip || len > cap - op) return -1;
        /* end-of-block rules, lz4.c:4104-4164 */
        int closing = (op + len > cap - 12) || (ip + len > clen - 8);
        if (closing && is_last && ip + len != clen) return -1;
        memcpy(dst + op, src + ip, (size_t)len); ip += len; op += len;
        if (closing && (is_last || op == cap)) break;
        if (ip == clen) break;                  /* partition ended on a literal run */
        if (ip + 2 > clen) return -1;
        int64_t off = src[ip] | (src[ip + 1] << 8); ip += 2;
        len = tok & 15;
        if (len == 15) {
            uint32_t b;
            do { if (ip >= clen) return -1; b = src[ip++]; len += b; } while (b == 255);
        }
        len += 4;
        if (off == 0 || off > op) return -1;    /* lz4.c:4196-4197 */
        if (len > cap - op) return -1;
        if (is_last && op + len > cap - 5) return -1;   /* lz4.c:4262-4264 */
        for (int64_t k = 0; k < len; k++) dst[op + k] = dst[op + k - off];
        op += len;
        if (!is_last && (op == cap || ip >= clen)) break;   /* lz4.c:4285-4288 */
    }
    return op;
}

int64_t orc_lz4_decompress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap)
{
    int64_t frame;
    if (src == NULL || dst == NULL) return -1;
    int T = rap_read_header(src, n, &frame);
    if (T < 0) return -1;
    if (T == 1) return orc_lz4_decode_partition(src + frame, n - frame, dst, cap, 1);
    int64_t total = 0;
    for (int i = 0; i < T; i++) {                /* lz4.c:4820-4881 */
        const uint8_t *e = src + 16 + 12 * (int64_t)i;
        uint32_t off = ld32(e), clen = ld32(e + 4), dlen = ld32(e + 8);
        if (clen == 0) continue;                 /* threads/threads.c:264-268 */
        if ((int64_t)off + clen > n || total + dlen > cap) return -1;
        int64_t got = orc_lz4_decode_partition(src + off, clen, dst + total, dlen, i == T - 1);
        if (got != (int64_t)dlen) return -1;
        total += got;
    }
    return total;
}

/* ---------------------------------------------------------------- Snappy encoder */
static uint8_t *put_varint32(uint8_t *p, uint32_t v)   /* snappy-stubs-internal.h:440-470 */
{
    while (v >= 128) { *p++ = (uint8_t)(v | 128); v >>= 7; }
    *p++ = (uint8_t)v;
    return p;
}
/* returns bytes consumed (0 on error) */
static int get_varint32(const uint8_t *p, int64_t n, uint32_t *out)
{
    uint32_t v = 0;
    for (int i = 0; i < 5; i++) {
        if (i >= n) return 0;
        uint32_t b = p[i];
        if (i == 4 && b > 15) return 0;          /* snappy-stubs-internal.h:418-437 */
        v |= (b & 127) << (7 * i);
        if (b < 128) { *out = v; return i + 1; }
    }
    return 0;
}
static uint8_t *snappy_put_literal(uint8_t *op, const uint8_t *lit, int64_t len)   /* snappy.cc:436-476 */
{
    uint32_t nm1 = (uint32_t)(len - 1);
    if (nm1 < 60) *op++ = (uint8_t)(nm1 << 2);
    else {
        int count = 1; for (uint32_t t = nm1 >> 8; t; t >>= 8) count++;
        *op++ = (uint8_t)((59 + count) << 2);
        for (int k = 0; k < count; k++) *op++ = (uint8_t)(nm1 >> (8 * k));
    }
    memcpy(op, lit, (size_t)len);
    return op + len;
}
static uint8_t *snappy_put_copy_le64(uint8_t *op, uint32_t off, uint32_t len)      /* snappy.cc:479-505 */
{
    if (len < 12 && off < 2048) {
        *op++ = (uint8_t)(1 | ((len - 4) << 2) | ((off >> 8) << 5));
        *op++ = (uint8_t)off;
    } else {
        *op++ = (uint8_t)(2 | ((len - 1) << 2));
        st16(op, off); op += 2;
    }
    return op;
}
static uint8_t *snappy_put_copy(uint8_t *op, uint32_t off, uint32_t len)           /* snappy.cc:540-568 */
{
    while (len >= 68) { op = snappy_put_copy_le64(op, off, 64); len -= 64; }
    if (len > 64) { op = snappy_put_copy_le64(op, off, 60); len -= 60; }
    return snappy_put_copy_le64(op, off, len);
}

int64_t orc_snappy_encode_fragment(const uint8_t *src, int64_t n, uint8_t *dst)
{
    /* table size, snappy.cc:619-657 */
    uint32_t tsize = 256;
    if (n > 16384) tsize = 16384; else while (tsize < (uint32_t)n) tsize <<= 1;
    int shift = 32; for (uint32_t t = tsize; t > 1; t >>= 1) shift--;
    uint16_t *tab = (uint16_t *)calloc(tsize, 2);
    uint8_t *op = dst;
    int64_t ip = 0;
#define SNAP_HASH(p) ((ld32(src + (p)) * 0x1e35a7bdU) >> shift)       /* snappy.cc:152-158 */
    if (n >= 15) {
        const int64_t ip_limit = n - 15;
        for (;;) {
            int64_t next_emit = ip++, cand;
            uint32_t skip = 32;
            for (;;) {                           /* snappy.cc:903-974; the 16-probe prologue is the */
                uint32_t stride = skip >> 5;     /* same walk with skip = 32..47, i.e. stride 1     */
                skip += stride;
                if (ip + stride > ip_limit) { ip = next_emit; goto remainder; }
                uint32_t h = SNAP_HASH(ip);
                cand = tab[h];
                tab[h] = (uint16_t)ip;
                if (ld32(src + cand) == ld32(src + ip)) break;
                ip += stride;
            }
            op = snappy_put_literal(op, src + next_emit, ip - next_emit);   /* snappy.cc:980 */
            for (;;) {                           /* snappy.cc:995-1032 */
                int64_t len = 4 + common_prefix(src + ip + 4, src + cand + 4, src + n);
                op = snappy_put_copy(op, (uint32_t)(ip - cand), (uint32_t)len);
                ip += len;
                if (ip >= ip_limit) goto remainder;
                tab[SNAP_HASH(ip - 1)] = (uint16_t)(ip - 1);
                uint32_t h = SNAP_HASH(ip);
                cand = tab[h];
                tab[h] = (uint16_t)ip;
                if (ld32(src + cand) != ld32(src + ip)) break;
            }
        }
    }
remainder:
    if (ip < n) op = snappy_put_literal(op, src + ip, n - ip);            /* snappy.cc:1039-1043 */
#undef SNAP_HASH
    free(tab);
    return op - dst;
}

/* one partition body = its fragments back to back (snappy.cc:1762-1818, minus the varint) */
static int64_t snappy_encode_body(const uint8_t *src, int64_t n, uint8_t *dst)
{
    uint8_t *op = dst;
    for (int64_t p = 0; p < n; p += 65536) {
        int64_t fn = n - p < 65536 ? n - p : 65536;
        op += orc_snappy_encode_fragment(src + p, fn, op);
    }
    return op - dst;
}

int64_t orc_snappy_compress(const uint8_t *src, int64_t n, uint8_t *dst, int max_threads)
{
    int T = orc_partition_count(n, ORC_SNAPPY_WINDOW, ORC_WINDOW_FACTOR, max_threads);
    uint8_t *op = dst;
    if (T == 1) {                                /* snappy.cc:2518-2527 */
        op = put_varint32(op, (uint32_t)n);
        op += snappy_encode_body(src, n, op);
        return op - dst;
    }
    int64_t common = n / T, left = n % T;
    int64_t frame = rap_write_header(dst, T);
    op = put_varint32(dst + frame, (uint32_t)n); /* snappy.cc:2617-2619 */
    for (int i = 0; i < T; i++) {                /* snappy.cc:2626-2651 */
        int64_t pn = common + (i == T - 1 ? left : 0);
        int64_t bn = snappy_encode_body(src + common * i, pn, op);
        rap_write_entry(dst, i, (uint32_t)(op - dst), (uint32_t)bn, (uint32_t)pn);
        op += bn;
    }
    return op - dst;
}

/* ---------------------------------------------------------------- Snappy decoder */
int64_t orc_snappy_decode_body(const uint8_t *src, int64_t clen, uint8_t *dst, int64_t expect)
{
    int64_t ip = 0, op = 0;
    while (ip < clen) {
        uint32_t tag = src[ip++];
        int64_t len, off;
        if ((tag & 3) == 0) {                    /* literal, snappy.cc:1492-1527 */
            len = (tag >> 2) + 1;
            if (len > 60) {
                int nb = (int)len - 60;
                if (ip + nb > clen) return -1;
                uint32_t v = 0;
                for (int k = 0; k < nb; k++) v |= (uint32_t)src[ip + k] << (8 * k);
                ip += nb;
                len = (int64_t)v + 1;
            }
            if (len > clen - ip || len > expect - op) return -1;
            memcpy(dst + op, src + ip, (size_t)len); ip += len; op += len;
            continue;
        }
        switch (tag & 3) {                       /* char_table, snappy-internal.h:406-439 */
        case 1:
            if (ip + 1 > clen) return -1;
            len = 4 + ((tag >> 2) & 7); off = ((tag >> 5) << 8) | src[ip]; ip += 1; break;
        case 2:
            if (ip + 2 > clen) return -1;
            len = 1 + (tag >> 2); off = src[ip] | (src[ip + 1] << 8); ip += 2; break;
        default:
            if (ip + 4 > clen) return -1;
            len = 1 + (tag >> 2); off = ld32(src + ip); ip += 4; break;
        }
        if (off == 0 || off > op || len > expect - op) return -1;   /* snappy.cc:2185-2199 */
        for (int64_t k = 0; k < len; k++) dst[op + k] = dst[op + k - off];
        op += len;
    }
    return op == expect ? op : -1;               /* snappy.cc:1715 */
}

int64_t orc_snappy_uncompressed_length(const uint8_t *src, int64_t n)
{
    int64_t frame; uint32_t v;
    if (src == NULL) return -1;
    int T = rap_read_header(src, n, &frame);
    if (T < 0) frame = 0;   /* snappy.cc:596-615 ignores the setup status; offset -1 is UB there */
    if (frame > n) return -1;
    if (!get_varint32(src + frame, n - frame, &v)) return -1;
    return v;
}

int64_t orc_snappy_decompress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap)
{
    int64_t frame; uint32_t total;
    if (src == NULL) return -1;
    int T = rap_read_header(src, n, &frame);
    if (T < 0 || frame > n) return -1;
    int vb = get_varint32(src + frame, n - frame, &total);
    if (!vb || (int64_t)total > cap) return -1;  /* api/codec.cpp:289 */
    if (T == 1)
        return orc_snappy_decode_body(src + frame + vb, n - frame - vb, dst, total);
    int64_t done = 0;
    for (int i = 0; i < T; i++) {                /* snappy.cc:2315-2366 */
        const uint8_t *e = src + 16 + 12 * (int64_t)i;
        uint32_t off = ld32(e), clen = ld32(e + 4), dlen = ld32(e + 8);
        if (clen == 0) continue;
        if ((int64_t)off + clen > n || done + dlen > cap) return -1;
        if (orc_snappy_decode_body(src + off, clen, dst + done, dlen) < 0) return -1;
        done += dlen;
    }
    return done == (int64_t)total ? done : -1;
}
