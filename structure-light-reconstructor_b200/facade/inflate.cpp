// inflate.cpp — zlib-stream (RFC 1950 / 1951) decoder for the PNG ingest path.
//
// cv::imread(path, 0) on 2 x 14 camera images is what MFReconstruct::loadCamImgs spends its time in
// (Duke/mfreconstruct.cpp:119-134): camera frames carry sensor noise, so their PNG streams are long runs of Huffman
// literals and the general-purpose inflate of zlib decodes them at ~120 MB/s per core — 11 ms per 1280x1024 frame,
// two hundred times the GPU time of the whole scan.  Entropy decoding of ONE deflate stream is sequential, so it stays
// on the host cores (one stream per thread), but it does not have to be slow.  This decoder is written for that
// input:
//   * 64-bit bit buffer refilled with one unaligned 8-byte load (>= 56 valid bits after every refill, enough for a
//     length + distance pair or three literals without another refill);
//   * one table lookup per symbol: 11-bit primary table for literals / lengths (8-bit for distances) whose entries
//     carry the decoded value, its extra-bit count and the code length, with second-level tables for longer codes;
//   * up to three literals per loop iteration, matches copied eight bytes at a time;
//   * the whole output buffer is known in advance (PNG: height x (1 + row bytes)), so there is no window and no
//     streaming state.
// Malformed input is rejected (over-subscribed code-length sets, undefined codes, distances before the start of the
// output, overruns of either buffer, wrong Adler-32), never trusted.  Checked against zlib on every compression level
// and strategy in tests/test_facade_host.py.
#include "inflate.h"

#include <string.h>

namespace duke {
namespace {

constexpr int LIT_TB = 12;    // primary table bits, literal / length codes
constexpr int DIST_TB = 8;    // primary table bits, distance codes
constexpr uint32_t F_LITERAL = 0x8000u, F_EOB = 0x4000u, F_SUB = 0x2000u, F_INVALID = 0x1000u, F_DOUBLE = 0x40u;
constexpr uint32_t LEN_MASK = 0x3fu;
// entry: bits 0..5 code bits to consume (second level: bits beyond the primary ones), bit 6 F_DOUBLE, 8..11 extra-bit
// count / second-level index bits, flags above, bits 16.. = literal value(s) / base length / base distance / index of
// the second-level table.  F_LITERAL | F_DOUBLE: TWO literals (bits 16..23 then 24..31) whose codes together fit the
// primary index — camera noise after PNG's Sub filter is a handful of small residuals with 3..5-bit codes, so most
// lookups of such a stream yield two output bytes.
constexpr int LIT_TABLE_MAX = (1 << LIT_TB) + 288 * 16;
constexpr int DIST_TABLE_MAX = (1 << DIST_TB) + 32 * 128;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t reverse_bits(uint32_t v, int n)   // the low n <= 16 bits of v, reversed
{
    v = ((v >> 1) & 0x5555u) | ((v & 0x5555u) << 1);
    v = ((v >> 2) & 0x3333u) | ((v & 0x3333u) << 2);
    v = ((v >> 4) & 0x0f0fu) | ((v & 0x0f0fu) << 4);
    v = ((v >> 8) & 0x00ffu) | ((v & 0x00ffu) << 8);
    return v >> (16 - n);
}

enum Kind { KIND_LITLEN, KIND_DIST, KIND_PRECODE };

inline uint32_t symbol_entry(Kind kind, int sym)
{
    if (kind == KIND_LITLEN) {
        if (sym < 256) return ((uint32_t)sym << 16) | F_LITERAL;
        if (sym == 256) return F_EOB;
        if (sym > 285) return F_INVALID;                      // 286, 287 take part in the code but never occur
        return ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_EXTRA[sym - 257] << 8);
    }
    if (kind == KIND_DIST) {
        if (sym > 29) return F_INVALID;
        return ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_EXTRA[sym] << 8);
    }
    return (uint32_t)sym << 16;
}

// Canonical Huffman decode table from code lengths.  Returns false for an over-subscribed set.  Incomplete sets are
// accepted (RFC 1951 allows a single distance code); their unused patterns decode to F_INVALID.
bool build_table(const uint8_t *lens, int n, Kind kind, int tb, uint32_t *table, int table_max)
{
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    uint32_t next_code[16] = {0};
    uint32_t code = 0;
    for (int l = 1; l <= 15; l++) {
        left = left * 2 - count[l];
        if (left < 0) return false;
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
    }
    const int primary = 1 << tb;
    if (left > 0)   // incomplete code: some patterns stay undefined (a complete code writes every primary entry below)
        for (int i = 0; i < primary; i++) table[i] = F_INVALID;
    uint32_t codes[288];
    int long_syms[288], n_long = 0;
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        codes[s] = reverse_bits(next_code[l]++, l);
        if (l > tb) long_syms[n_long++] = s;
    }
    if (n_long) {
        // second-level tables, one per primary pattern that longer codes start with, sized by the longest of them
        uint8_t submax[1 << LIT_TB];
        memset(submax, 0, (size_t)primary);
        for (int k = 0; k < n_long; k++) {
            const int s = long_syms[k];
            uint8_t &m = submax[codes[s] & (uint32_t)(primary - 1)];
            if (lens[s] - tb > m) m = (uint8_t)(lens[s] - tb);
        }
        int used = primary;
        for (int k = 0; k < n_long; k++) {
            const uint32_t pfx = codes[long_syms[k]] & (uint32_t)(primary - 1);
            if (!submax[pfx]) continue;                       // allocated by an earlier symbol
            const int size = 1 << submax[pfx];
            if (used + size > table_max) return false;
            table[pfx] = ((uint32_t)used << 16) | F_SUB | ((uint32_t)submax[pfx] << 8) | (uint32_t)tb;
            for (int j = 0; j < size; j++) table[used + j] = F_INVALID;
            used += size;
            submax[pfx] = 0;
        }
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t ent = symbol_entry(kind, s);
        if (l <= tb) {
            for (uint32_t k = codes[s]; k < (uint32_t)primary; k += 1u << l) table[k] = ent | (uint32_t)l;
        } else {
            const uint32_t p = table[codes[s] & (uint32_t)(primary - 1)];
            const int sub_bits = (int)((p >> 8) & 0xf), base = (int)(p >> 16);
            for (uint32_t k = codes[s] >> tb; k < (1u << sub_bits); k += 1u << (l - tb)) table[base + k] = ent | (uint32_t)(l - tb);
        }
    }
    if (kind == KIND_LITLEN) {
        // pair up literals: index i = [code 1][code 2][..]; the bits left after code 1 decide the second symbol only if
        // its code is no longer than what is left.  Descending order: entry i >> l1 <= i is still a single symbol.
        for (int i = primary - 1; i >= 0; i--) {
            const uint32_t e1 = table[i];
            if (!(e1 & F_LITERAL)) continue;
            const int l1 = (int)(e1 & LEN_MASK);
            const uint32_t e2 = table[i >> l1];
            const int l2 = (int)(e2 & LEN_MASK);
            if ((e2 & F_LITERAL) && l1 + l2 <= tb)
                table[i] = (e1 & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8) | F_LITERAL | F_DOUBLE | (uint32_t)(l1 + l2);
        }
    }
    return true;
}

struct Tables {
    uint32_t lit[LIT_TABLE_MAX];
    uint32_t dist[DIST_TABLE_MAX];
};

struct BitReader {
    const uint8_t *ip, *iend;
    uint64_t buf = 0;
    int cnt = 0;          // valid bits in buf
    size_t overrun = 0;   // zero bytes supplied beyond the end of the input

    inline void refill()
    {
        if (iend - ip >= 8) {
            uint64_t w;
            memcpy(&w, ip, 8);   // little endian host (x86-64, aarch64)
            buf |= w << cnt;
            ip += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt < 56) {
                if (ip < iend)
                    buf |= (uint64_t)*ip++ << cnt;
                else
                    overrun++;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void consume(int n)
    {
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t take(int n)
    {
        const uint32_t v = peek(n);
        consume(n);
        return v;
    }
    // more bits consumed than the input held?
    inline bool past_end() const { return overrun * 8 > (size_t)cnt; }
};

inline uint32_t lookup(const uint32_t *table, int tb, BitReader &br)
{
    uint32_t e = table[br.peek(tb)];
    if (e & F_SUB) {
        br.consume(tb);
        e = table[(e >> 16) + br.peek((int)((e >> 8) & 0xf))];
    }
    return e;
}

bool fail(std::string *err, const char *m)
{
    if (err) *err = m;
    return false;
}

const Tables *static_tables()
{
    static Tables t;
    static bool built = [] {
        uint8_t l[288], d[32];
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        for (int i = 0; i < 32; i++) d[i] = 5;
        return build_table(l, 288, KIND_LITLEN, LIT_TB, t.lit, LIT_TABLE_MAX) &&
               build_table(d, 32, KIND_DIST, DIST_TB, t.dist, DIST_TABLE_MAX);
    }();
    return built ? &t : nullptr;
}

inline void copy_match(uint8_t *op, size_t dist, size_t len)
{
    const uint8_t *src = op - dist;
    if (dist >= 8) {   // 8 bytes at a time; may write up to 7 bytes past op + len (the caller left room)
        uint8_t *end = op + len;
        do {
            uint64_t w;
            memcpy(&w, src, 8);
            memcpy(op, &w, 8);
            src += 8;
            op += 8;
        } while (op < end);
    } else if (dist == 1) {
        memset(op, *src, len);
    } else {
        for (size_t k = 0; k < len; k++) op[k] = src[k];
    }
}

// One Huffman-coded block.  The fast loop runs while both buffers have slack for a worst-case iteration (three
// literals or a 258-byte match with its 8-byte copy granularity; 8 input bytes per refill); the careful loop finishes.
bool inflate_block(BitReader &br_io, const Tables &t, uint8_t *out, uint8_t *&op_io, uint8_t *oend, std::string *err)
{
    // the reader and the output pointer live in locals: byte stores may alias anything the compiler cannot prove
    // private, and a bit buffer that is spilled around every literal costs more than the decoding itself
    BitReader br = br_io;
    uint8_t *op = op_io;
    const uint32_t *const lit = t.lit;
    const uint32_t lit_mask = (1u << LIT_TB) - 1;
    uint32_t e;
    for (;;) {
        // ---- fast loop: room for a worst-case iteration in both buffers (six literals or a 258-byte match copied in
        // 8-byte steps; one 8-byte load per refill) ----
        while ((oend - op) >= 320 && (br.iend - br.ip) >= 16) {
            {
                uint64_t w;
                memcpy(&w, br.ip, 8);
                br.buf |= w << br.cnt;
                br.ip += (63 - br.cnt) >> 3;
                br.cnt |= 56;
            }
            e = lit[br.buf & lit_mask];
            // up to three lookups on one refill (3 x 15 bits <= 56); an entry may carry two literals: both bytes are
            // stored, the pointer moves by one or two
#define DUKE_PUT_LITERALS(e)                                   \
    do {                                                       \
        const uint16_t two = (uint16_t)((e) >> 16);            \
        br.buf >>= ((e) & LEN_MASK);                           \
        br.cnt -= (int)((e) & LEN_MASK);                       \
        memcpy(op, &two, 2);                                   \
        op += 1 + (((e) >> 6) & 1u);                           \
    } while (0)
            if (e & F_LITERAL) {
                DUKE_PUT_LITERALS(e);
                e = lit[br.buf & lit_mask];
                if (e & F_LITERAL) {
                    DUKE_PUT_LITERALS(e);
                    e = lit[br.buf & lit_mask];
                    if (e & F_LITERAL) {
                        DUKE_PUT_LITERALS(e);
                        continue;
                    }
                }
                br.refill();
            }
#undef DUKE_PUT_LITERALS
            if (e & F_SUB) {
                br.consume(LIT_TB);
                e = lit[(e >> 16) + br.peek((int)((e >> 8) & 0xf))];
                if (e & F_LITERAL) {
                    br.consume((int)(e & LEN_MASK));
                    *op++ = (uint8_t)(e >> 16);
                    continue;
                }
            }
            if (e & (F_EOB | F_INVALID)) goto block_end;
            {
                // length + distance: <= 15 + 5 + 15 + 13 = 48 bits, all in the buffer
                br.consume((int)(e & LEN_MASK));
                const size_t len = (e >> 16) + br.take((int)((e >> 8) & 0xf));
                const uint32_t d = lookup(t.dist, DIST_TB, br);
                if (d & F_INVALID) return fail(err, "inflate: undefined distance code");
                br.consume((int)(d & LEN_MASK));
                const size_t dist = (d >> 16) + br.take((int)((d >> 8) & 0xf));
                if (dist > (size_t)(op - out)) return fail(err, "inflate: distance reaches before the start of the output");
                copy_match(op, dist, len);
                op += len;
            }
        }
        // ---- careful step: every access checked ----
        br.refill();
        e = lit[br.buf & lit_mask];
        if (e & F_SUB) {
            br.consume(LIT_TB);
            e = lit[(e >> 16) + br.peek((int)((e >> 8) & 0xf))];
        }
        if (e & F_LITERAL) {
            const size_t nlit = 1 + ((e >> 6) & 1u);
            if ((size_t)(oend - op) < nlit) return fail(err, "inflate: output overrun");
            br.consume((int)(e & LEN_MASK));
            op[0] = (uint8_t)(e >> 16);
            if (nlit == 2) op[1] = (uint8_t)(e >> 24);
            op += nlit;
            if (br.past_end()) return fail(err, "inflate: input ends inside a block");
            continue;
        }
        if (e & (F_EOB | F_INVALID)) goto block_end;
        {
            br.consume((int)(e & LEN_MASK));
            const size_t len = (e >> 16) + br.take((int)((e >> 8) & 0xf));
            const uint32_t d = lookup(t.dist, DIST_TB, br);
            if (d & F_INVALID) return fail(err, "inflate: undefined distance code");
            br.consume((int)(d & LEN_MASK));
            const size_t dist = (d >> 16) + br.take((int)((d >> 8) & 0xf));
            if (dist > (size_t)(op - out)) return fail(err, "inflate: distance reaches before the start of the output");
            if (len > (size_t)(oend - op)) return fail(err, "inflate: output overrun");
            const uint8_t *src = op - dist;
            for (size_t k = 0; k < len; k++) op[k] = src[k];
            op += len;
            if (br.past_end()) return fail(err, "inflate: input ends inside a block");
        }
    }
block_end:
    if (e & F_INVALID) return fail(err, "inflate: undefined literal/length code");
    br.consume((int)(e & LEN_MASK));
    if (br.past_end()) return fail(err, "inflate: input ends inside a block");
    br_io = br;
    op_io = op;
    return true;
}

bool read_dynamic_tables(BitReader &br, Tables &t, std::string *err)
{
    br.refill();
    const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return fail(err, "inflate: bad dynamic block header");
    uint8_t pre_lens[19] = {0};
    for (int i = 0; i < hclen; i++) {
        br.refill();
        pre_lens[PRECODE_ORDER[i]] = (uint8_t)br.take(3);
    }
    uint32_t pre[1 << 7];
    if (!build_table(pre_lens, 19, KIND_PRECODE, 7, pre, 1 << 7)) return fail(err, "inflate: bad code-length code");
    uint8_t lens[286 + 30 + 140];
    int n = 0;
    while (n < hlit + hdist) {
        br.refill();
        const uint32_t e = pre[br.peek(7)];
        if (e & F_INVALID) return fail(err, "inflate: undefined code-length code");
        br.consume((int)(e & LEN_MASK));
        const int sym = (int)(e >> 16);
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
        } else {
            int rep;
            uint8_t v = 0;
            if (sym == 16) {
                if (n == 0) return fail(err, "inflate: repeat with no previous code length");
                v = lens[n - 1];
                rep = 3 + (int)br.take(2);
            } else if (sym == 17) {
                rep = 3 + (int)br.take(3);
            } else {
                rep = 11 + (int)br.take(7);
            }
            if (n + rep > hlit + hdist) return fail(err, "inflate: code lengths overrun the header counts");
            memset(lens + n, v, (size_t)rep);
            n += rep;
        }
        if (br.past_end()) return fail(err, "inflate: input ends inside a block header");
    }
    if (lens[256] == 0) return fail(err, "inflate: no end-of-block code");
    if (!build_table(lens, hlit, KIND_LITLEN, LIT_TB, t.lit, LIT_TABLE_MAX)) return fail(err, "inflate: over-subscribed literal/length code");
    if (!build_table(lens + hlit, hdist, KIND_DIST, DIST_TB, t.dist, DIST_TABLE_MAX)) return fail(err, "inflate: over-subscribed distance code");
    return true;
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target_clones("avx2", "default")))   // the sums below vectorise 3x better with 256-bit integer lanes
#endif
uint32_t adler32(const uint8_t *p, size_t n)
{
    // a' = a + sum p[i], b' = b + m*a + sum (m - i) p[i] over a chunk of m bytes: two independent sums the compiler
    // vectorises, instead of the textbook two-adds-per-byte dependency chain
    uint32_t a = 1, b = 0;
    constexpr size_t CHUNK = 256;   // 256 * 255 * 256 < 2^32
    while (n >= CHUNK) {
        uint32_t s1 = 0, s2 = 0;
        for (size_t i = 0; i < CHUNK; i++) {
            s1 += p[i];
            s2 += (uint32_t)(CHUNK - i) * p[i];
        }
        b = (b + (uint32_t)CHUNK * a + s2) % 65521u;
        a = (a + s1) % 65521u;
        p += CHUNK;
        n -= CHUNK;
    }
    while (n--) {
        a += *p++;
        b += a;
    }
    return ((b % 65521u) << 16) | (a % 65521u);
}

}  // namespace

bool raw_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, size_t *in_used, std::string *err)
{
    BitReader br;
    br.ip = in;
    br.iend = in + in_len;
    uint8_t *op = out, *oend = out + out_len;
    Tables *dyn = nullptr;
    bool ok = true;
    for (;;) {
        br.refill();
        const uint32_t final_block = br.take(1), type = br.take(2);
        if (type == 0) {
            br.consume(br.cnt & 7);                            // to the byte boundary
            br.refill();
            const uint32_t len = br.take(16), nlen = br.take(16);
            if ((len ^ nlen) != 0xffffu || br.past_end()) { ok = fail(err, "inflate: bad stored block"); break; }
            // bytes still in the bit buffer belong to the block: step the input pointer back to them
            const uint8_t *src = br.ip - (br.cnt >> 3) + br.overrun;
            if ((size_t)(br.iend - src) < len || br.overrun) { ok = fail(err, "inflate: stored block overruns the input"); break; }
            if ((size_t)(oend - op) < len) { ok = fail(err, "inflate: output overrun"); break; }
            memcpy(op, src, len);
            op += len;
            br.ip = src + len;
            br.buf = 0;
            br.cnt = 0;
        } else if (type == 1) {
            const Tables *st = static_tables();
            if (!st || !inflate_block(br, *st, out, op, oend, err)) { ok = false; break; }
        } else if (type == 2) {
            if (!dyn) dyn = new Tables;
            if (!read_dynamic_tables(br, *dyn, err) || !inflate_block(br, *dyn, out, op, oend, err)) { ok = false; break; }
        } else {
            ok = fail(err, "inflate: reserved block type");
            break;
        }
        if (final_block) break;
    }
    delete dyn;
    if (!ok) return false;
    if (op != oend) return fail(err, "inflate: stream is shorter than the expected output");
    if (in_used) *in_used = (size_t)(br.ip - in) - (size_t)(br.cnt >> 3) + br.overrun;
    return true;
}

bool zlib_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, std::string *err)
{
    if (in_len < 6) return fail(err, "zlib: stream too short");
    const unsigned cmf = in[0], flg = in[1];
    if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0) return fail(err, "zlib: bad header");
    if (flg & 0x20) return fail(err, "zlib: preset dictionary");
    size_t used = 0;
    if (!raw_inflate(in + 2, in_len - 2, out, out_len, &used, err)) return false;
    if (used + 4 > in_len - 2) return fail(err, "zlib: missing Adler-32");
    const uint8_t *t = in + 2 + used;
    const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
    if (adler32(out, out_len) != want) return fail(err, "zlib: Adler-32 mismatch");
    return true;
}

}  // namespace duke
