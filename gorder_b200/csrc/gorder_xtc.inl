// gorder_xtc.inl — host-side trajectory feed of the engine: GROMACS XTC frames are decoded
// by a pool of host threads straight into pinned batches in the engine's plane layout and handed to
// gorder_gpu_submit_native; the decode of batch k+1 overlaps the H2D copy and the kernels of batch k.
//
// Stands in for the reference's reader when the harness (tests, bench.py) needs a real end-to-end number:
//   read_trajectory -> groan_rs traj_iter_map_reduce::<GroupXtcReader> (src/analysis/common.rs:281-304) -> molly 0.5.0
// (Cargo.lock:955, not vendored).  The Rust host keeps its own reader; this file only has to produce
// the same coordinates: integer lattice * (1 / precision) in f32, as xdrfile's xdr3dfcoord and molly do.
// Format: big-endian XDR header (magic 1995 | 2023, natoms, step, time, box[9], natoms, precision, minint[3],
// maxint[3], smallidx, byte count) followed by the "xtc3" bit stream: per atom a mixed-radix triple of
// sizeint[] in `bitsize` bits, a run flag (+ 5 bits), then run/3 triples relative to the previous atom in `smallidx`
// bits with the first pair swapped (water trick); smallidx walks through magicints[] as the runs ask.
// Included at the end of gorder_capi.cu.
#include <atomic>
#include <chrono>
#include <climits>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <type_traits>
#include <unistd.h>

namespace gxtc {

const int kMagic[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                      645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768,
                      41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255,
                      1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085,
                      16777216};
constexpr int kFirstIdx = 9, kLastIdx = (int)(sizeof(kMagic) / sizeof(kMagic[0])) - 1;

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline float bef32(const uint8_t *p) { uint32_t u = be32(p); float f; memcpy(&f, &u, 4); return f; }

struct Frame {
    size_t payload;      // offset of the compressed bytes (or of the raw floats when natoms <= 9)
    size_t nbytes;
    int natoms, step, smallidx;
    float time, precision, box[9];
    int minint[3], maxint[3];
};

// MSB-first bit reader over [p, end): one unaligned big-endian 64-bit load per field while at least 16 bytes remain,
// byte-wise near the end of the stream; reads past `end` yield zeros
struct BitIn {
    const uint8_t *p, *end;
    uint64_t pos = 0;   // bit offset from p
    inline bool roomy() const { return (pos >> 3) + 16 <= (uint64_t)(end - p); }
    inline uint64_t window() const {   // the >= 57 bits that follow pos, left-aligned
        uint64_t w;
        memcpy(&w, p + (pos >> 3), 8);
        return __builtin_bswap64(w) << (pos & 7);
    }
    inline uint32_t get_slow(int bits) {
        uint64_t v = 0;
        for (int i = 0; i < bits; i++) {
            const uint64_t at = pos + (uint64_t)i;
            const uint8_t *b = p + (at >> 3);
            v = (v << 1) | (uint64_t)(b < end ? (*b >> (7 - (at & 7))) & 1 : 0);
        }
        pos += (uint64_t)bits;
        return (uint32_t)v;
    }
    // SURE: the caller has checked that the whole group lies at least 16 bytes before the end of the stream
    template <bool SURE = false> inline uint32_t get(int bits) {   // bits <= 32
        if (bits == 0) return 0;
        if (!SURE && !roomy()) return get_slow(bits);
        const uint32_t v = (uint32_t)(window() >> (64 - bits));
        pos += (uint64_t)bits;
        return v;
    }
    // the mixed-radix number of a triple: its bytes arrive least significant first, the last one may be partial
    template <typename W, bool SURE = false> inline W get_le(int bits) {
        if (sizeof(W) == 8 && bits <= 56 && bits > 0 && (SURE || roomy())) {
            const uint64_t w = window();
            const int full = bits >> 3, rem = bits & 7;
            uint64_t v = full ? (__builtin_bswap64(w) & (full == 8 ? ~0ull : ((1ull << (8 * full)) - 1ull))) : 0ull;
            if (rem) v |= ((w << (8 * full)) >> (64 - rem)) << (8 * full);
            pos += (uint64_t)bits;
            return (W)v;
        }
        W v = 0;
        int shift = 0;
        while (bits > 8) { v |= (W)get<SURE>(8) << shift; shift += 8; bits -= 8; }
        if (bits > 0) v |= (W)get<SURE>(bits) << shift;
        return v;
    }
};

inline int bits_of(unsigned size) { unsigned num = 1; int b = 0; while (size >= num && b < 32) { b++; num <<= 1; } return b; }
inline int bits_of_triple(const unsigned s[3]) {   // xdrfile's sizeofints: the bit length of s0 * s1 * s2
    unsigned __int128 t = (unsigned __int128)s[0] * s[1] * s[2];
    int b = 0;
    while (t) { b++; t >>= 1; }
    return b;
}

template <typename W>
inline void unpack3(W v, const unsigned s[3], int out[3]) {
    out[2] = (int)(v % s[2]); v /= s[2];
    out[1] = (int)(v % s[1]); v /= s[1];
    out[0] = (int)v;
}

// v / d by one 64 x 64 -> 128 bit multiplication with m = ceil(2^64 / d): exact whenever v * d < 2^64
struct FastDiv {
    uint64_t m = 0;
    unsigned d = 1;
    bool ok = false;
    inline void set(unsigned div, int value_bits) {   // values are below 2^value_bits
        d = div ? div : 1;
        int db = 0;
        while (db < 32 && (1ull << db) <= d) db++;
        ok = d > 1 && value_bits + db < 64;
        m = ok ? (uint64_t)((((unsigned __int128)1 << 64) + d - 1) / d) : 0;
    }
    inline uint64_t div(uint64_t v) const { return ok ? (uint64_t)(((unsigned __int128)v * m) >> 64) : v / d; }
};
inline void unpack3_fast(uint64_t v, const FastDiv &d1, const FastDiv &d2, int out[3]) {
    const uint64_t q = d2.div(v);
    out[2] = (int)(v - q * d2.d);
    const uint64_t r = d1.div(q);
    out[1] = (int)(q - r * d1.d);
    out[0] = (int)r;
}

struct SmallDivs {   // the divisor of the small triples for every smallidx (the number has smallidx bits)
    FastDiv d[sizeof(kMagic) / sizeof(kMagic[0])];
    SmallDivs() { for (int i = 0; i <= kLastIdx; i++) d[i].set((unsigned)kMagic[i], i); }
};

// Decode one frame; emit(atom, x, y, z) is called for atoms 0 .. stop_after (inclusive) in order.
template <typename Emit>
int decode_frame(const uint8_t *base, const Frame &fr, int stop_after, Emit emit) {
    const int n = fr.natoms;
    const int last = std::min(stop_after, n - 1);
    if (n <= 9) {
        for (int i = 0; i <= last; i++) emit(i, bef32(base + fr.payload + 12 * (size_t)i), bef32(base + fr.payload + 12 * (size_t)i + 4), bef32(base + fr.payload + 12 * (size_t)i + 8));
        return 0;
    }
    unsigned sizeint[3], bitsint[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) sizeint[k] = (unsigned)(fr.maxint[k] - fr.minint[k] + 1);
    int bitsize = 0;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) { for (int k = 0; k < 3; k++) bitsint[k] = (unsigned)bits_of(sizeint[k]); }
    else bitsize = bits_of_triple(sizeint);
    int smallidx = fr.smallidx;
    if (smallidx < kFirstIdx || smallidx > kLastIdx) return -1;
    int smaller = kMagic[std::max(kFirstIdx, smallidx - 1)] / 2, smallnum = kMagic[smallidx] / 2;
    unsigned ssz[3] = {(unsigned)kMagic[smallidx], (unsigned)kMagic[smallidx], (unsigned)kMagic[smallidx]};
    const float inv = 1.0f / fr.precision;
    BitIn in{base + fr.payload, base + fr.payload + fr.nbytes};
    static const SmallDivs kSmall;
    FastDiv big1, big2;   // divisors of the large triple (per frame)
    big1.set(sizeint[1], bitsize); big2.set(sizeint[2], bitsize);
    const FastDiv *small = &kSmall.d[smallidx];
    int i = 0, run = 0, cur[3], prev[3];
    // one group (a "large" atom and the run of small ones behind it): at most 128 + 6 + 10 * 72 bits = 107 bytes, even in a corrupt stream
    auto group = [&](auto sure_tag) -> int {
        constexpr bool SURE = decltype(sure_tag)::value;
        if (bitsize == 0) { cur[0] = (int)in.template get<SURE>((int)bitsint[0]); cur[1] = (int)in.template get<SURE>((int)bitsint[1]); cur[2] = (int)in.template get<SURE>((int)bitsint[2]); }
        else if (bitsize <= 64) unpack3_fast(in.template get_le<uint64_t, SURE>(bitsize), big1, big2, cur);
        else unpack3<unsigned __int128>(in.template get_le<unsigned __int128, SURE>(bitsize), sizeint, cur);
        cur[0] += fr.minint[0]; cur[1] += fr.minint[1]; cur[2] += fr.minint[2];
        prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
        int is_smaller = 0;
        bool flag;
        if (SURE || in.roomy()) {   // flag and run length in one look
            const uint32_t six = (uint32_t)(in.window() >> 58);
            flag = (six & 32u) != 0;
            if (flag) run = (int)(six & 31u);
            in.pos += flag ? 6 : 1;
        } else {
            flag = in.get(1) != 0;
            if (flag) run = (int)in.get(5);
        }
        if (flag) {
            is_smaller = run % 3;
            run -= is_smaller;
            is_smaller--;
        }
        if (run > 0) {
            for (int k = 0; k < run; k += 3) {
                int d[3];
                if (smallidx <= 64) unpack3_fast(in.template get_le<uint64_t, SURE>(smallidx), *small, *small, d);
                else unpack3<unsigned __int128>(in.template get_le<unsigned __int128, SURE>(smallidx), ssz, d);
                int t[3] = {d[0] + prev[0] - smallnum, d[1] + prev[1] - smallnum, d[2] + prev[2] - smallnum};
                if (i + (k == 0 ? 2 : 1) > n) return -1;
                if (k == 0) {   // the first two atoms of a run are stored in swapped order: this one comes first ...
                    emit(i, (float)t[0] * inv, (float)t[1] * inv, (float)t[2] * inv); i++;
                    emit(i, (float)prev[0] * inv, (float)prev[1] * inv, (float)prev[2] * inv); i++;   // ... then the "large" atom
                } else {
                    emit(i, (float)t[0] * inv, (float)t[1] * inv, (float)t[2] * inv); i++;
                }
                prev[0] = t[0]; prev[1] = t[1]; prev[2] = t[2];   // the run continues relative to the atom just decoded
            }
        } else {
            if (i >= n) return -1;
            emit(i, (float)cur[0] * inv, (float)cur[1] * inv, (float)cur[2] * inv); i++;
        }
        smallidx += is_smaller;
        if (smallidx < kFirstIdx || smallidx > kLastIdx) return -1;
        if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > kFirstIdx ? kMagic[smallidx - 1] / 2 : 0; }
        else if (is_smaller > 0) { smaller = smallnum; smallnum = kMagic[smallidx] / 2; }
        if (is_smaller) { ssz[0] = ssz[1] = ssz[2] = (unsigned)kMagic[smallidx]; small = &kSmall.d[smallidx]; }
        return 0;
    };
    while (i <= last) {
        const int rc = (in.pos >> 3) + 128 <= (uint64_t)fr.nbytes ? group(std::true_type{}) : group(std::false_type{});
        if (rc) return rc;
    }
    return 0;
}

// ---- writer (synthetic trajectories for the bench and the tests) -------------------------------------------
struct BitOut {
    std::vector<uint8_t> buf;
    uint64_t acc = 0;
    int n = 0;
    inline void put(uint32_t v, int bits) {
        if (bits == 0) return;
        acc = (acc << bits) | (v & ((bits == 32) ? 0xffffffffull : ((1ull << bits) - 1ull)));
        n += bits;
        while (n >= 8) { n -= 8; buf.push_back((uint8_t)(acc >> n)); }
    }
    template <typename W> inline void put_le(W v, int bits) {
        while (bits > 8) { put((uint32_t)(v & 0xff), 8); v >>= 8; bits -= 8; }
        if (bits > 0) put((uint32_t)v, bits);
    }
    inline void flush() { if (n > 0) { buf.push_back((uint8_t)(acc << (8 - n))); n = 0; } }
};
inline void wr32(std::vector<uint8_t> &o, uint32_t v) { o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v); }
inline void wrf(std::vector<uint8_t> &o, float f) { uint32_t u; memcpy(&u, &f, 4); wr32(o, u); }

template <typename W> inline W pack3(const unsigned s[3], const int v[3]) { return ((W)(unsigned)v[0] * s[1] + (unsigned)v[1]) * s[2] + (unsigned)v[2]; }
inline void put_triple(BitOut &out, int bits, const unsigned s[3], const int v[3]) {
    if (bits <= 64) out.put_le<uint64_t>(pack3<uint64_t>(s, v), bits);
    else out.put_le<unsigned __int128>(pack3<unsigned __int128>(s, v), bits);
}

// xdrfile's compression strategy (runs of atoms within magicints[smallidx] / 2 of their predecessor)
inline void encode_frame(std::vector<uint8_t> &o, const float *xyz, const float *box3, int natoms, int step, float time, float precision) {
    wr32(o, 1995); wr32(o, (uint32_t)natoms); wr32(o, (uint32_t)step); wrf(o, time);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) wrf(o, r == c ? box3[r] : 0.0f);
    wr32(o, (uint32_t)natoms);
    if (natoms <= 9) { for (int i = 0; i < 3 * natoms; i++) wrf(o, xyz[i]); return; }
    wrf(o, precision);
    std::vector<int> L(3 * (size_t)natoms);
    int mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    std::vector<int> diffs;
    diffs.reserve((size_t)natoms);
    for (int i = 0; i < natoms; i++) {
        for (int k = 0; k < 3; k++) {
            const float lf = xyz[3 * (size_t)i + k] * precision;
            const int li = (int)(lf >= 0.0f ? lf + 0.5f : lf - 0.5f);
            L[3 * (size_t)i + k] = li;
            mn[k] = std::min(mn[k], li); mx[k] = std::max(mx[k], li);
        }
        if (i > 0) {
            long long d = 0;
            for (int k = 0; k < 3; k++) d += std::llabs((long long)L[3 * (size_t)i + k] - L[3 * (size_t)(i - 1) + k]);
            diffs.push_back((int)std::min<long long>(d, INT32_MAX));
        }
    }
    // xdrfile starts smallidx from the MINIMUM displacement between consecutive atoms.  In a real trajectory sterics keep
    // that minimum at ~0.1 nm; in a synthetic one a single accidental close pair among a million atoms would switch the
    // runs off for the whole frame, so the 1st percentile is used instead (any starting smallidx is a valid stream).
    long long mindiff = INT32_MAX;
    if (!diffs.empty()) {
        const size_t q = diffs.size() / 100;
        std::nth_element(diffs.begin(), diffs.begin() + q, diffs.end());
        mindiff = diffs[q];
    }
    for (int k = 0; k < 3; k++) wr32(o, (uint32_t)mn[k]);
    for (int k = 0; k < 3; k++) wr32(o, (uint32_t)mx[k]);
    unsigned sizeint[3], bitsint[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) sizeint[k] = (unsigned)(mx[k] - mn[k] + 1);
    int bitsize = 0;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) { for (int k = 0; k < 3; k++) bitsint[k] = (unsigned)bits_of(sizeint[k]); }
    else bitsize = bits_of_triple(sizeint);
    int smallidx = kFirstIdx;
    while (smallidx < kLastIdx && kMagic[smallidx] < mindiff) smallidx++;
    wr32(o, (uint32_t)smallidx);
    const int maxidx = std::min(kLastIdx, smallidx + 8), minidx = maxidx - 8;
    int smaller = kMagic[std::max(kFirstIdx, smallidx - 1)] / 2, smallnum = kMagic[smallidx] / 2;
    const int larger = kMagic[maxidx] / 2;
    unsigned ssz[3] = {(unsigned)kMagic[smallidx], (unsigned)kMagic[smallidx], (unsigned)kMagic[smallidx]};
    BitOut out;
    out.buf.reserve((size_t)natoms * 5);
    int prevrun = -1, i = 0, prev[3] = {0, 0, 0};
    auto near = [&](const int *a, const int *b, int lim) { return std::abs(a[0] - b[0]) < lim && std::abs(a[1] - b[1]) < lim && std::abs(a[2] - b[2]) < lim; };
    while (i < natoms) {
        int *cur = &L[3 * (size_t)i];
        int is_small = 0, is_smaller;
        if (smallidx < maxidx && i >= 1 && near(cur, prev, larger)) is_smaller = 1;
        else if (smallidx > minidx) is_smaller = -1;
        else is_smaller = 0;
        if (i + 1 < natoms && near(cur, cur + 3, smallnum)) {
            for (int k = 0; k < 3; k++) std::swap(cur[k], cur[3 + k]);   // store the pair in swapped order (undone by the reader)
            is_small = 1;
        }
        int tmp[3] = {cur[0] - mn[0], cur[1] - mn[1], cur[2] - mn[2]};
        if (bitsize == 0) { out.put((uint32_t)tmp[0], (int)bitsint[0]); out.put((uint32_t)tmp[1], (int)bitsint[1]); out.put((uint32_t)tmp[2], (int)bitsint[2]); }
        else put_triple(out, bitsize, sizeint, tmp);
        prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
        i++;
        int run = 0, small[24];
        if (is_small == 0 && is_smaller == -1) is_smaller = 0;
        while (is_small && run < 8 * 3) {
            const int *nx = &L[3 * (size_t)i];
            if (is_smaller == -1) {
                const long long dx = nx[0] - prev[0], dy = nx[1] - prev[1], dz = nx[2] - prev[2];
                if (dx * dx + dy * dy + dz * dz >= (long long)smaller * smaller) is_smaller = 0;
            }
            small[run++] = nx[0] - prev[0] + smallnum; small[run++] = nx[1] - prev[1] + smallnum; small[run++] = nx[2] - prev[2] + smallnum;
            prev[0] = nx[0]; prev[1] = nx[1]; prev[2] = nx[2];
            i++;
            is_small = (i < natoms && near(&L[3 * (size_t)i], prev, smallnum)) ? 1 : 0;
        }
        if (run != prevrun || is_smaller != 0) { prevrun = run; out.put(1, 1); out.put((uint32_t)(run + is_smaller + 1), 5); }
        else out.put(0, 1);
        for (int k = 0; k < run; k += 3) put_triple(out, smallidx, ssz, small + k);
        if (is_smaller != 0) {
            smallidx += is_smaller;
            if (is_smaller < 0) { smallnum = smaller; smaller = kMagic[smallidx - 1] / 2; }
            else { smaller = smallnum; smallnum = kMagic[smallidx] / 2; }
            ssz[0] = ssz[1] = ssz[2] = (unsigned)kMagic[smallidx];
        }
    }
    out.flush();
    wr32(o, (uint32_t)out.buf.size());
    o.insert(o.end(), out.buf.begin(), out.buf.end());
    while (o.size() & 3) o.push_back(0);
}

}  // namespace gxtc

struct GorderXtc {
    int fd = -1;
    const uint8_t *data = nullptr;
    size_t size = 0;
    std::vector<gxtc::Frame> frames;
    int natoms = 0;
};

extern "C" {

void gorder_xtc_close(GorderXtc *x) {
    if (!x) return;
    if (x->data) munmap((void *)x->data, x->size);
    if (x->fd >= 0) close(x->fd);
    delete x;
}

int gorder_xtc_open(const char *path, GorderXtc **out) {
    if (!path || !out) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    GorderXtc *x = new GorderXtc();
    x->fd = open(path, O_RDONLY);
    struct stat st;
    if (x->fd < 0 || fstat(x->fd, &st) != 0 || st.st_size < 60) { gorder_xtc_close(x); return GORDER_ERR_INVALID_ARGUMENT; }
    x->size = (size_t)st.st_size;
    void *m = mmap(nullptr, x->size, PROT_READ, MAP_PRIVATE, x->fd, 0);
    if (m == MAP_FAILED) { x->data = nullptr; gorder_xtc_close(x); return GORDER_ERR_INVALID_ARGUMENT; }
    x->data = (const uint8_t *)m;
    size_t p = 0;
    while (p + 60 <= x->size) {   // index the frames: headers only
        const uint8_t *d = x->data + p;
        const uint32_t magic = gxtc::be32(d);
        if (magic != 1995 && magic != 2023) { gorder_xtc_close(x); return GORDER_ERR_INVALID_ARGUMENT; }
        gxtc::Frame f{};
        f.natoms = (int)gxtc::be32(d + 4); f.step = (int)gxtc::be32(d + 8); f.time = gxtc::bef32(d + 12);
        for (int i = 0; i < 9; i++) f.box[i] = gxtc::bef32(d + 16 + 4 * i);
        if ((int)gxtc::be32(d + 52) != f.natoms || f.natoms <= 0) { gorder_xtc_close(x); return GORDER_ERR_INVALID_ARGUMENT; }
        size_t q = p + 56;
        if (f.natoms <= 9) { f.payload = q; f.nbytes = 12 * (size_t)f.natoms; f.precision = 0.0f; q += f.nbytes; }
        else {
            if (q + 36 > x->size) break;
            f.precision = gxtc::bef32(x->data + q);
            for (int k = 0; k < 3; k++) { f.minint[k] = (int)gxtc::be32(x->data + q + 4 + 4 * k); f.maxint[k] = (int)gxtc::be32(x->data + q + 16 + 4 * k); }
            f.smallidx = (int)gxtc::be32(x->data + q + 28);
            q += 32;
            if (magic == 2023) { f.nbytes = ((size_t)gxtc::be32(x->data + q) << 32) | gxtc::be32(x->data + q + 4); q += 8; }
            else { f.nbytes = gxtc::be32(x->data + q); q += 4; }
            f.payload = q;
            // a forged / corrupt length (the 64-bit field of magic 2023 can hold anything) must not wrap the cursor
            if (f.nbytes > x->size - q) break;   // truncated (or hostile) frame: indexing stops here
            q += std::min((f.nbytes + 3) & ~(size_t)3, x->size - q);   // the padding of the last frame may be missing
        }
        if (q > x->size || q <= p) break;   // truncated last frame: ignored, as trajectory readers do
        if (x->frames.empty()) x->natoms = f.natoms;
        else if (f.natoms != x->natoms) { gorder_xtc_close(x); return GORDER_ERR_INVALID_ARGUMENT; }
        x->frames.push_back(f);
        p = q;
    }
    if (x->frames.empty()) { gorder_xtc_close(x); return GORDER_ERR_INVALID_ARGUMENT; }
    *out = x;
    return GORDER_OK;
}

int gorder_xtc_info(GorderXtc *x, int32_t *n_atoms, int64_t *n_frames, float *precision) {
    if (!x) return GORDER_ERR_INVALID_ARGUMENT;
    if (n_atoms) *n_atoms = x->natoms;
    if (n_frames) *n_frames = (int64_t)x->frames.size();
    if (precision) *precision = x->frames[0].precision;
    return GORDER_OK;
}

// Decode frames first, first + stride, ... (count of them) into host arrays: xyz [count][n_atoms][3], box9 [count][9],
// time [count], step [count] (any of the last three may be NULL).
int gorder_xtc_read(GorderXtc *x, int64_t first, int64_t count, int64_t stride, int32_t n_threads, float *xyz, float *box9, float *time, int32_t *step) {
    if (!x || !xyz || first < 0 || count < 0 || stride < 1 || (count > 0 && first + (count - 1) * stride >= (int64_t)x->frames.size())) return GORDER_ERR_INVALID_ARGUMENT;
    std::atomic<int64_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        for (;;) {
            const int64_t j = next.fetch_add(1);
            if (j >= count) break;
            const gxtc::Frame &f = x->frames[(size_t)(first + j * stride)];
            float *dst = xyz + (size_t)j * x->natoms * 3;
            if (gxtc::decode_frame(x->data, f, x->natoms - 1, [&](int i, float a, float b, float c) { dst[3 * (size_t)i] = a; dst[3 * (size_t)i + 1] = b; dst[3 * (size_t)i + 2] = c; })) bad = 1;
            if (box9) memcpy(box9 + 9 * j, f.box, sizeof(f.box));
            if (time) time[j] = f.time;
            if (step) step[j] = f.step;
        }
    };
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, count));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return bad ? GORDER_ERR_INVALID_ARGUMENT : GORDER_OK;
}

// Write (append = 0: create) frames with orthogonal boxes: xyz [n_frames][n_atoms][3], box3 [n_frames][3]; frame f gets
// step = first_step + f and time = step * dt.
int gorder_xtc_write(const char *path, const float *xyz, const float *box3, int32_t n_atoms, int64_t n_frames, float precision, int32_t append,
                     int32_t first_step, float dt, int32_t n_threads) {
    if (!path || !xyz || !box3 || n_atoms <= 0 || n_frames < 0 || !(precision > 0.0f)) return GORDER_ERR_INVALID_ARGUMENT;
    std::vector<std::vector<uint8_t>> enc((size_t)n_frames);
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t f = next.fetch_add(1);
            if (f >= n_frames) break;
            gxtc::encode_frame(enc[(size_t)f], xyz + (size_t)f * n_atoms * 3, box3 + 3 * f, n_atoms, first_step + (int)f, (first_step + (int)f) * dt, precision);
        }
    };
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_frames));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    FILE *fp = fopen(path, append ? "ab" : "wb");
    if (!fp) return GORDER_ERR_INVALID_ARGUMENT;
    for (auto &e : enc) if (fwrite(e.data(), 1, e.size(), fp) != e.size()) { fclose(fp); return GORDER_ERR_INVALID_ARGUMENT; }
    fclose(fp);
    return GORDER_OK;
}

// analyze_frame for the frames first, first + stride, ... < last of an open trajectory: n_threads host threads decode
// straight into pinned batches in the plane layout (only the atoms the engine needs; the bit stream of a frame is read
// up to the last of them), gorder_gpu_submit_native runs them.  atom_of_slot[s] = trajectory atom of engine atom s
// (NULL: identity).  frame_index of the j-th analysed frame is frame_index0 + j * stride (topology/mod.rs:141-144;
// frame_index0 = 0, or the continuation value when several files are concatenated, common.rs:306-339).
// decode_seconds (optional): host time spent decoding, summed over threads.
int gorder_gpu_run_xtc(GorderHandle *h, GorderXtc *x, const int32_t *atom_of_slot, int64_t first, int64_t last, int64_t stride, int64_t frame_index0,
                       int32_t n_threads, int32_t batch_frames, double *decode_seconds) {
    if (!h || !x || first < 0 || stride < 1) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (h->err_code) return h->err_code;
    last = std::min<int64_t>(last, (int64_t)x->frames.size());
    const int64_t total = last > first ? (last - first + stride - 1) / stride : 0;
    if (total == 0) return GORDER_OK;
    cudaSetDevice(h->device);
    const int na = h->s.n_atoms;
    // trajectory atom -> (plane offset, component stride); several engine atoms may not share a trajectory atom
    std::vector<int> off_of_atom((size_t)x->natoms, -1), cs_of_atom((size_t)x->natoms, 0);
    int stop_after = -1;
    for (int s = 0; s < na; s++) {
        const int a = atom_of_slot ? atom_of_slot[s] : s;
        if (a < 0 || a >= x->natoms) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "atom_of_slot outside the trajectory", a); return h->err_code; }
        if (h->slot_off[s] < 0) continue;   // not needed by the analysis
        off_of_atom[a] = h->slot_off[s]; cs_of_atom[a] = h->slot_cs[s];
        stop_after = std::max(stop_after, a);
    }
    const int B = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(batch_frames > 0 ? batch_frames : 32, h->max_batch), total));
    const size_t ff = (size_t)h->frame_floats;
    // pinned batch buffers live in the handle (pinning hundreds of MB costs more than decoding them)
    if (h->xtc_pin_frames < B) {
        for (int i = 0; i < 2; i++) { if (h->xtc_pin[i]) cudaFreeHost(h->xtc_pin[i]); if (h->xtc_pbox[i]) cudaFreeHost(h->xtc_pbox[i]); h->xtc_pin[i] = h->xtc_pbox[i] = nullptr; }
        h->xtc_pin_frames = 0;
        for (int i = 0; i < 2; i++) {
            if (cudaMallocHost((void **)&h->xtc_pin[i], (size_t)B * ff * sizeof(float)) != cudaSuccess ||
                cudaMallocHost((void **)&h->xtc_pbox[i], (size_t)B * 3 * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                h->set_error(GORDER_ERR_OUT_OF_MEMORY, "pinned batch buffers"); return h->err_code;
            }
            memset(h->xtc_pin[i], 0, (size_t)B * ff * sizeof(float));   // padding stays finite
        }
        h->xtc_pin_frames = B;
    }
    float *pin[2] = {h->xtc_pin[0], h->xtc_pin[1]}, *pbox[2] = {h->xtc_pbox[0], h->xtc_pbox[1]};
    std::atomic<int> bad{0};
    std::atomic<long long> dec_ns{0};
    auto decode_batch = [&](int64_t j0, int nf, int buf) {
        std::atomic<int> next{0};
        auto work = [&]() {
            const auto t0 = std::chrono::steady_clock::now();
            for (;;) {
                const int j = next.fetch_add(1);
                if (j >= nf) break;
                const gxtc::Frame &f = x->frames[(size_t)(first + (j0 + j) * stride)];
                // check_box (common.rs:186-198): orthogonal boxes only
                if (f.box[1] != 0.0f || f.box[2] != 0.0f || f.box[3] != 0.0f || f.box[5] != 0.0f || f.box[6] != 0.0f || f.box[7] != 0.0f) bad = GORDER_ERR_NOT_ORTHOGONAL_BOX;
                pbox[buf][3 * j] = f.box[0]; pbox[buf][3 * j + 1] = f.box[4]; pbox[buf][3 * j + 2] = f.box[8];
                float *dst = pin[buf] + (size_t)j * ff;
                const int *oo = off_of_atom.data(), *cc = cs_of_atom.data();
                if (gxtc::decode_frame(x->data, f, stop_after, [&](int i, float a, float b, float c) {
                        const int o = oo[i];
                        if (o >= 0) { const size_t cs = (size_t)cc[i]; dst[o] = a; dst[o + cs] = b; dst[o + 2 * cs] = c; }
                    })) bad = GORDER_ERR_INVALID_ARGUMENT;
            }
            dec_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        };
        const int nt = std::max(1, std::min<int>(n_threads, nf));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    };
    int rc = GORDER_OK;
    std::vector<int64_t> fi((size_t)B);
    int64_t j0 = 0;
    int nf = (int)std::min<int64_t>(B, total);
    decode_batch(0, nf, 0);
    for (int buf = 0; j0 < total && !rc; buf ^= 1) {
        if (bad) { h->set_error(bad, bad == GORDER_ERR_NOT_ORTHOGONAL_BOX ? "simulation box is not orthogonal" : "corrupt XTC frame"); rc = h->err_code; break; }
        for (int j = 0; j < nf; j++) fi[(size_t)j] = frame_index0 + (j0 + j) * stride;
        // the next batch is decoded by a helper thread (which runs the pool) while this one is copied and analysed
        const int64_t j1 = j0 + nf;
        const int nf1 = (int)std::min<int64_t>(B, total - j1);
        std::thread ahead;
        if (nf1 > 0) ahead = std::thread([&, j1, nf1, buf]() { decode_batch(j1, nf1, buf ^ 1); });
        rc = gorder_gpu_submit_native(h, pin[buf], pbox[buf], fi.data(), nf);
        if (ahead.joinable()) ahead.join();
        j0 = j1; nf = nf1;
    }
    if (decode_seconds) *decode_seconds = (double)dec_ns.load() * 1e-9;
    return rc;
}

}  // extern "C"

// =====================================================================================================================
// Device-side XTC unpacking ("later a GPU XTC bit-unpacker"): the GPU does the arithmetic of the
// decode, the host only copies the compressed frames and bookmarks them.  From decoded floats the end-to-end path moves
// 12 B per atom over PCIe, from a file the host decode bounds it; the compressed stream is 4-6 B per
// atom.
//
// The bit stream of a frame is sequential only in its CONTROL information: where a group (one "large" atom + run / 3
// "small" atoms) starts depends on the run flags and on smallidx, not on the coordinates.
//   host (the threads that copy a frame into the pinned batch): walk the control bits only -- 6 bits looked at per
//        group, no arithmetic -- and drop a bookmark (bit offset, first atom, run, smallidx) every 32 groups;
//   xtc_decode_kernel: one thread per group.  Lane j of a warp starts at the warp's bookmark, walks j groups forward
//        (control bits only), then extracts the mixed-radix triples of ITS group (64-bit arithmetic), rebuilds the
//        lattice points of its atoms and writes  int * (1 / precision)  in f32 -- the reader's arithmetic -- into the
//        [atom][xyz] staging frame of the engine (atoms the analysis does not need are dropped).
// (A first version walked the whole frame on the device, one lane per frame: 133 ms per frame-walk whatever the batch,
//  i.e. ~1000 frames in flight to keep PCIe busy; the bookmarks cost the host ~1 ms per frame and thread.)
// Frames the device path does not cover (more than 64 bits per triple, <= 9 atoms) take the host decoder.
// =====================================================================================================================
namespace gxtc {

constexpr int kBookmarkEvery = 32;   // groups per bookmark = lanes per warp
struct Bookmark { unsigned pos, atom0; unsigned short run, sidx; unsigned pad; };

struct DevFrame {
    unsigned long long payload;   // byte offset of the frame's stream in the batch buffer (multiple of 16)
    unsigned long long bookmarks; // index of the frame's first bookmark
    unsigned nbytes;
    int natoms, n_groups, bitsize;
    int minint[3];
    unsigned sizeint[3];
    int bitsint[3];
    float inv_precision;
    int smallidx;                 // of the frame header: where the walk starts
    int pad_;
};

// host: bookmarks of one frame; returns the number of groups, -1 when the stream is inconsistent, -2 when it needs
// more than 64 bits per small triple (not covered by the device path).
// Branch-free in the data: the run code is translated by two 32-entry tables (code = run + is_smaller + 1).
struct RunTables {
    signed char smalls[32], delta[32];
    RunTables() { for (int c = 0; c < 32; c++) { const int sm = c % 3; smalls[c] = (signed char)((c - sm) / 3); delta[c] = (signed char)(sm - 1); } }
};
// The walk over the control bits of one frame, one group per step() so that two frames can be walked in the same loop:
// a step is a chain of dependent operations (look at 6 bits, two table look-ups, a multiply-add: ~12 ns), two independent
// chains overlap in the core's out-of-order window.
struct Walker {
    const uint8_t *p = nullptr;
    size_t nbytes = 0;
    unsigned long long end = 0, last_safe = 0, pos = 0;
    int i = 0, g = 0, smalls = 0, sidx = 0, n = 0, large_bits = 0;
    int rc = 0;   // 0: walking, 1: finished, -1: inconsistent stream, -2: > 64 bits per small triple
    std::vector<Bookmark> *out = nullptr;
    inline void init(const uint8_t *base, const Frame &f, int large_bits_, std::vector<Bookmark> *out_) {
        p = base + f.payload; nbytes = f.nbytes; end = (unsigned long long)f.nbytes * 8ull; pos = 0;
        i = 0; g = 0; smalls = 0; sidx = f.smallidx; n = f.natoms; large_bits = large_bits_; out = out_;
        rc = 0;
        if (end < 16) { rc = -1; return; }
        last_safe = end - 16;   // the two-byte look stays inside the stream below this offset
        if (n <= 0) rc = 1;
    }
    inline void step(const RunTables &T) {   // rc == 0
        if (sidx > 64) { rc = -2; return; }
        if (sidx < kFirstIdx) { rc = -1; return; }
        if ((g % kBookmarkEvery) == 0) out->push_back(Bookmark{(unsigned)pos, (unsigned)i, (unsigned short)(3 * smalls), (unsigned short)sidx, 0u});
        pos += (unsigned long long)large_bits;
        const size_t b = (size_t)(pos >> 3);
        unsigned two;
        if (pos > last_safe) {   // tail of the stream: careful byte access
            if (pos + 1 > end) { rc = -1; return; }
            two = ((unsigned)p[b] << 8) | (b + 1 < nbytes ? p[b + 1] : 0u);
        } else two = ((unsigned)p[b] << 8) | p[b + 1];
        const unsigned six = (two >> (10 - (pos & 7))) & 63u;
        const unsigned flag = six >> 5, code = six & 31u;
        smalls = flag ? T.smalls[code] : smalls;
        pos += 1 + 5 * flag + (unsigned long long)(smalls * sidx);
        i += 1 + smalls;
        sidx += flag ? T.delta[code] : 0;
        g++;
        if (i >= n) rc = (i != n || pos > end) ? -1 : 1;
    }
    inline int result() const { return rc == 1 ? g : rc; }
};
inline int bookmark_frame(const uint8_t *base, const Frame &f, int large_bits, std::vector<Bookmark> &out) {
    static const RunTables T;
    Walker w;
    w.init(base, f, large_bits, &out);
    while (w.rc == 0) w.step(T);
    return w.result();
}
// page cache -> pinned batch with non-temporal stores (dst 16-byte aligned): the pinned lines are read next by the copy
// engine, not by a core, so they need not be fetched for ownership nor kept in the cache: 76 instead of 49 GB/s with 16 threads
// on the bench box, whose memory system is what bounds this path (GORDER_XTC_NO_NT_COPY: plain memcpy)
inline void stream_copy(unsigned char *dst, const unsigned char *src, size_t n) {
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i)), b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32)), d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a); _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c); _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
        }
    memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}
// up to kWalkWays frames at once (see Walker): ng[k] = groups of frame k or its error code
constexpr int kWalkWays = 4;
inline void bookmark_many(const uint8_t *base, const Frame *const *fr, const int *bits, std::vector<Bookmark> *const *out, int *ng, int count) {
    static const RunTables T;
    Walker w[kWalkWays];
    for (int k = 0; k < count; k++) w[k].init(base, *fr[k], bits[k], out[k]);
    for (;;) {
        int active = 0;
        for (int k = 0; k < count; k++)
            if (w[k].rc == 0) { w[k].step(T); active++; }
        if (!active) break;
    }
    for (int k = 0; k < count; k++) ng[k] = w[k].result();
}

__constant__ int c_magic[73];

// n <= 32 bits starting at bit `pos` of a big-endian bit stream stored in 32-bit words
__device__ __forceinline__ unsigned dev_bits(const unsigned *__restrict__ w, unsigned pos, int n) {
    if (n == 0) return 0u;
    const unsigned i = pos >> 5;
    const unsigned w0 = __byte_perm(__ldg(w + i), 0, 0x0123), w1 = __byte_perm(__ldg(w + i + 1), 0, 0x0123);
    const unsigned v = __funnelshift_l(w1, w0, pos & 31u);
    return v >> (32 - n);
}
// the mixed-radix number of a triple (n <= 64): bytes least significant first, the last one may be partial
__device__ __forceinline__ unsigned long long dev_le(const unsigned *__restrict__ w, unsigned pos, int n) {
    unsigned long long v = 0;
    int shift = 0;
    while (n >= 32) {   // four whole bytes: the stream's first byte is the least significant one
        v |= (unsigned long long)__byte_perm(dev_bits(w, pos, 32), 0, 0x0123) << shift;
        pos += 32; shift += 32; n -= 32;
    }
    if (n > 0) {
        const unsigned b = dev_bits(w, pos, n);
        int left = n;
        while (left > 8) { left -= 8; v |= (unsigned long long)((b >> left) & 0xffu) << shift; shift += 8; }
        v |= (unsigned long long)(b & ((1u << left) - 1u)) << shift;
    }
    return v;
}
__device__ __forceinline__ void dev_unpack3(unsigned long long v, unsigned s1, unsigned s2, int (&out)[3]) {
    if (v >> 32) {
        const unsigned long long q = v / s2; out[2] = (int)(v - q * s2);
        const unsigned long long p = q / s1; out[1] = (int)(q - p * s1); out[0] = (int)p;
    } else {
        const unsigned x = (unsigned)v, q = x / s2; out[2] = (int)(x - q * s2);
        const unsigned p = q / s1; out[1] = (int)(q - p * s1); out[0] = (int)p;
    }
}

// one group (a "large" atom + its run of small ones): from the warp's bookmark to the group by the control bits, then unpack
__device__ __forceinline__ void xtc_decode_group(const unsigned char *__restrict__ bytes, const DevFrame &fr, const Bookmark *__restrict__ bookmarks,
                                                 const int *__restrict__ slot_of_atom, int n_engine_atoms, float *__restrict__ xyz, int f, int g) {
    const unsigned *w = reinterpret_cast<const unsigned *>(bytes + __ldg(&fr.payload));
    const int bitsize = __ldg(&fr.bitsize);
    const int large_bits = bitsize ? bitsize : __ldg(&fr.bitsint[0]) + __ldg(&fr.bitsint[1]) + __ldg(&fr.bitsint[2]);
    // ---- from the warp's bookmark to this lane's group: control bits only ----
    const Bookmark bm = bookmarks[__ldg(&fr.bookmarks) + (unsigned)(g / kBookmarkEvery)];
    unsigned pos = bm.pos;
    int i = (int)bm.atom0, run = bm.run, sidx = bm.sidx;
    for (int step = g % kBookmarkEvery; step > 0; step--) {
        pos += large_bits;
        const unsigned six = dev_bits(w, pos, 6);
        int is_smaller = 0;
        if (six & 32u) {
            const int code = (int)(six & 31u);
            is_smaller = code % 3;
            run = code - is_smaller;
            is_smaller--;
            pos += 6;
        } else pos += 1;
        const int smalls = run / 3;
        pos += (unsigned)(smalls * sidx);
        i += 1 + smalls;
        sidx += is_smaller;
    }
    // ---- this group ----
    const float inv = __ldg(&fr.inv_precision);
    float *out = xyz + (size_t)f * n_engine_atoms * 3;
    auto emit = [&](int atom, const int (&c)[3]) {
        const int s = __ldg(slot_of_atom + atom);
        if (s >= 0) { float *o = out + 3 * (size_t)s; o[0] = (float)c[0] * inv; o[1] = (float)c[1] * inv; o[2] = (float)c[2] * inv; }
    };
    int cur[3];
    if (bitsize) {
        dev_unpack3(dev_le(w, pos, bitsize), __ldg(&fr.sizeint[1]), __ldg(&fr.sizeint[2]), cur);
        pos += bitsize;
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { const int b = __ldg(&fr.bitsint[k]); cur[k] = (int)dev_bits(w, pos, b); pos += b; }
    }
    cur[0] += __ldg(&fr.minint[0]); cur[1] += __ldg(&fr.minint[1]); cur[2] += __ldg(&fr.minint[2]);
    {   // this group's own flag / run code
        const unsigned six = dev_bits(w, pos, 6);
        if (six & 32u) { const int code = (int)(six & 31u); run = code - code % 3; pos += 6; }
        else pos += 1;
    }
    if (run == 0) { emit(i, cur); return; }
    const unsigned ss = (unsigned)c_magic[sidx];
    const int smallnum = c_magic[sidx] / 2;
    int prev[3] = {cur[0], cur[1], cur[2]};
    for (int k = 0; k < run; k += 3) {
        int d[3];
        dev_unpack3(dev_le(w, pos, sidx), ss, ss, d);
        pos += sidx;
        const int t[3] = {d[0] + prev[0] - smallnum, d[1] + prev[1] - smallnum, d[2] + prev[2] - smallnum};
        if (k == 0) { emit(i, t); emit(i + 1, cur); i += 2; }   // the first two atoms of a run are stored in swapped order
        else { emit(i, t); i++; }
        prev[0] = t[0]; prev[1] = t[1]; prev[2] = t[2];
    }
}

// grid (x, frames); the x dimension strides over the frame's groups, whose number the device walk only knows on the device
__global__ void __launch_bounds__(256) xtc_decode_kernel(const unsigned char *__restrict__ bytes, const DevFrame *__restrict__ frames,
                                                         const Bookmark *__restrict__ bookmarks, const int *__restrict__ slot_of_atom,
                                                         int n_engine_atoms, float *__restrict__ xyz) {
    const int f = blockIdx.y;
    const DevFrame &fr = frames[f];
    const int ng = __ldg(&fr.n_groups);   // <= 0: nothing to do (empty, inconsistent or unsupported frame)
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += gridDim.x * blockDim.x)
        xtc_decode_group(bytes, fr, bookmarks, slot_of_atom, n_engine_atoms, xyz, f, g);
}

// The walk over the control bits on the DEVICE: one warp per frame; lane 0 follows the chain
//     position -> 6 bits of the stream -> run / smallidx -> next position
// and drops a bookmark every 32 groups -- the same walk as gxtc::Walker on the host (identical bookmarks) -- while all lanes
// stream the frame through a ring of four 4 KB chunks in shared memory with cp.async, two chunks ahead of the walker: a walk
// straight from global memory paid a DRAM miss for every new 32-byte sector (250 ns per group, 45 ms per 1 M-atom frame);
// from shared memory a step is ~70 cycles of dependent instructions.  That is still slow for ONE frame (~7 ms) and
// irrelevant for a batch: its frames walk side by side on as many SMs while the copy engine brings in the next batch, so the
// host has nothing left to do per frame but hand over the compressed bytes.  n_groups of the frame = its groups, or -1
// (inconsistent stream) / -2 (a small triple of more than 64 bits: not covered) -- the decoder skips such a frame and the
// engine's deferred error word takes GORDER_ERR_INVALID_ARGUMENT with the frame's number (reported by the next sync / finish).
constexpr int kWalkChunk = 4096, kWalkSlots = 4;
__global__ void __launch_bounds__(32) xtc_walk_kernel(const unsigned char *__restrict__ bytes, DevFrame *__restrict__ frames,
                                                      Bookmark *__restrict__ bookmarks, unsigned marks_per_frame, unsigned frame_cap,
                                                      int *__restrict__ err, long long *__restrict__ err_detail, long long frame0, long long frame_stride) {
    __shared__ __align__(16) unsigned ring[kWalkSlots * kWalkChunk / 4];
    const int f = blockIdx.x, lane = threadIdx.x;
    DevFrame &fr = frames[f];
    const unsigned char *src = bytes + fr.payload;   // 16-byte aligned
    const int n = fr.natoms;
    const int large_bits = fr.bitsize ? fr.bitsize : fr.bitsint[0] + fr.bitsint[1] + fr.bitsint[2];
    const unsigned end = fr.nbytes * 8u;             // nbytes < 2^29 (checked on the host)
    const unsigned limit = min(frame_cap, (fr.nbytes + 32u + 15u) & ~15u);   // bytes of the slot that hold the stream and its zeros
    const int n_chunks = (int)((fr.nbytes + 16u + kWalkChunk - 1) / kWalkChunk);
    Bookmark *bm = bookmarks + fr.bookmarks;
    auto load_chunk = [&](int c) {   // all lanes: chunk c -> its slot of the ring (zeros past the stream)
        if (c < n_chunks + 1) {
            const unsigned base = (unsigned)c * kWalkChunk;
            unsigned char *dst = reinterpret_cast<unsigned char *>(ring) + (size_t)(c % kWalkSlots) * kWalkChunk;
            for (unsigned k = (unsigned)lane * 16u; k < (unsigned)kWalkChunk; k += 512u) {
                if (base + k + 16u <= limit) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + k);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + base + k) : "memory");
                } else *reinterpret_cast<uint4 *>(dst + k) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_chunk(0);
    load_chunk(1);
    unsigned pos = 0;
    int i = 0, g = 0, smalls = 0, sidx = fr.smallidx, rc = 0;
    bool done = n <= 0 || sidx > 64 || sidx < kFirstIdx;
    if (n > 0 && sidx > 64) rc = -2; else if (n > 0 && sidx < kFirstIdx) rc = -1;
    for (int c = 0; !done; c++) {
        load_chunk(c + 2);
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // chunks <= c + 1 have landed
        __syncwarp();
        if (lane == 0) {
            // groups whose control bits START in chunk c (the look reaches 8 bytes further); in the last chunks `end` closes the window
            const unsigned stop = min((unsigned)(c + 1) * kWalkChunk * 8u, end);
            constexpr unsigned kRing = kWalkSlots * kWalkChunk / 4;
            for (;;) {   // one group per turn; the chain  pos -> 6 bits -> run, smallidx -> pos  is all there is: keep it short
                const unsigned cpos = pos + (unsigned)large_bits;
                if (cpos >= stop) { if (stop == end) rc = -1; break; }   // next chunk; or atoms left and no stream
                if ((g & (kBookmarkEvery - 1)) == 0) {
                    if ((unsigned)(g / kBookmarkEvery) >= marks_per_frame) { rc = -1; break; }
                    bm[g / kBookmarkEvery] = Bookmark{pos, (unsigned)i, (unsigned short)(3 * smalls), (unsigned short)sidx, 0u};
                }
                const unsigned wi = cpos >> 5;
                const unsigned w0 = __byte_perm(ring[wi % kRing], 0, 0x0123), w1 = __byte_perm(ring[(wi + 1) % kRing], 0, 0x0123);
                const unsigned six = __funnelshift_l(w1, w0, cpos & 31u) >> 26;
                pos = cpos + 1u;
                int run_bits = smalls * sidx;
                if (six & 32u) {
                    const int code = (int)(six & 31u);
                    smalls = (code * 11) >> 5;            // code / 3 for code < 32
                    run_bits = smalls * sidx;             // the run is read with the smallidx BEFORE its change
                    sidx += code - 3 * smalls - 1;
                    pos += 5u;
                }
                pos += (unsigned)run_bits;
                i += 1 + smalls;
                g++;
                if (i >= n) { rc = (i != n || pos > end) ? -1 : g; break; }
                if ((unsigned)(sidx - kFirstIdx) > (unsigned)(64 - kFirstIdx)) { rc = sidx > 64 ? -2 : -1; break; }   // as the host: before the next group
            }
        }
        rc = __shfl_sync(0xffffffffu, rc, 0);
        done = rc != 0;
        __syncwarp();
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (lane == 0) {
        if (n <= 0) rc = 0;
        fr.n_groups = rc;
        if (rc < 0 && atomicCAS(err, 0, (int)GORDER_ERR_INVALID_ARGUMENT) == 0) *err_detail = frame0 + (long long)f * frame_stride;
    }
}

}  // namespace gxtc

struct GorderXtcDev {   // device / pinned buffers of the unpacker, owned by the handle
    unsigned char *h_bytes[2] = {nullptr, nullptr}, *d_bytes[2] = {nullptr, nullptr};
    gxtc::DevFrame *h_frames[2] = {nullptr, nullptr}, *d_frames[2] = {nullptr, nullptr};
    gxtc::Bookmark *h_marks[2] = {nullptr, nullptr}, *d_marks[2] = {nullptr, nullptr};
    float *h_box[2] = {nullptr, nullptr}, *d_box[2] = {nullptr, nullptr};
    size_t bytes_cap = 0, marks_per_frame = 0, frame_cap = 0;
    int frames_cap = 0, n_traj_atoms = 0;
    int *d_slot_of_atom = nullptr;
    float *d_xyz = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    bool magic_uploaded = false;
    std::vector<int> slot_cache;   // host copy of d_slot_of_atom
    void free_all() {
        for (int i = 0; i < 2; i++) {
            if (h_bytes[i]) cudaFreeHost(h_bytes[i]);
            if (h_frames[i]) cudaFreeHost(h_frames[i]);
            if (h_marks[i]) cudaFreeHost(h_marks[i]);
            if (h_box[i]) cudaFreeHost(h_box[i]);
            cudaFree(d_bytes[i]); cudaFree(d_frames[i]); cudaFree(d_marks[i]); cudaFree(d_box[i]);
            if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
        }
        cudaFree(d_slot_of_atom); cudaFree(d_xyz);
    }
};

void gorder_xtc_dev_free(GorderXtcDev *d) { if (d) { d->free_all(); delete d; } }

extern "C" {

// The host stage of the device-decode path without a GPU: walk the control bits of frames first .. first + count - 1 (one
// look per group, two frames per loop as the staging threads do) and report, per frame, its groups and the bookmarks the
// kernel would get.  Cheap integrity check (the coordinates are not decoded).  n_groups[k] = -1: inconsistent stream (the call
// then returns GORDER_ERR_INVALID_ARGUMENT after scanning everything), -2: consistent, but a small triple needs more than 64
// bits (such frames take the host decoder); frames of <= 9 atoms are stored uncompressed: 0 groups.
int gorder_xtc_scan(GorderXtc *x, int64_t first, int64_t count, int32_t *n_groups, int32_t *n_bookmarks) {
    if (!x || first < 0 || count < 0 || first + count > (int64_t)x->frames.size()) return GORDER_ERR_INVALID_ARGUMENT;
    auto bits_of_frame = [&](const gxtc::Frame &f) {
        unsigned sz[3];
        for (int k = 0; k < 3; k++) sz[k] = (unsigned)(f.maxint[k] - f.minint[k] + 1);
        return (sz[0] | sz[1] | sz[2]) > 0xffffffu ? gxtc::bits_of(sz[0]) + gxtc::bits_of(sz[1]) + gxtc::bits_of(sz[2]) : gxtc::bits_of_triple(sz);
    };
    std::vector<gxtc::Bookmark> marks[gxtc::kWalkWays];
    int bad = 0;
    for (int64_t k0 = 0; k0 < count; k0 += gxtc::kWalkWays) {
        const int m = (int)std::min<int64_t>(gxtc::kWalkWays, count - k0);
        const gxtc::Frame *fr[gxtc::kWalkWays];
        std::vector<gxtc::Bookmark> *out[gxtc::kWalkWays];
        int bits[gxtc::kWalkWays], ng[gxtc::kWalkWays];
        for (int k = 0; k < m; k++) { fr[k] = &x->frames[(size_t)(first + k0 + k)]; bits[k] = bits_of_frame(*fr[k]); marks[k].clear(); out[k] = &marks[k]; }
        gxtc::bookmark_many(x->data, fr, bits, out, ng, m);
        for (int k = 0; k < m; k++) {
            int g = ng[k];
            size_t nm = marks[k].size();
            if (fr[k]->natoms <= 9) { g = 0; nm = 0; }
            if (g == -1) bad = 1;
            if (n_groups) n_groups[k0 + k] = g;
            if (n_bookmarks) n_bookmarks[k0 + k] = g < 0 ? 0 : (int32_t)nm;
        }
    }
    return bad ? GORDER_ERR_INVALID_ARGUMENT : GORDER_OK;
}

// gorder_gpu_run_xtc with the decode on the device.  Same arguments and results (the decoded coordinates are bit-identical
// to the host decoder's); n_threads host threads copy the compressed frames into the pinned batch and bookmark them.
// bytes_h2d (optional): bytes that crossed PCIe.
int gorder_gpu_run_xtc_device(GorderHandle *h, GorderXtc *x, const int32_t *atom_of_slot, int64_t first, int64_t last, int64_t stride,
                              int64_t frame_index0, int32_t n_threads, int32_t batch_frames, int64_t *bytes_h2d) {
    if (!h || !x || first < 0 || stride < 1) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (h->err_code) return h->err_code;
    last = std::min<int64_t>(last, (int64_t)x->frames.size());
    const int64_t total = last > first ? (last - first + stride - 1) / stride : 0;
    if (total == 0) return GORDER_OK;
    // frames the device path does not cover -> host decoder for the whole call
    size_t max_bytes = 0;
    for (int64_t j = 0; j < total; j++) {
        const gxtc::Frame &f = x->frames[(size_t)(first + j * stride)];
        unsigned sz[3];
        for (int k = 0; k < 3; k++) sz[k] = (unsigned)(f.maxint[k] - f.minint[k] + 1);
        const bool separate = (sz[0] | sz[1] | sz[2]) > 0xffffffu;
        if (f.natoms <= 9 || (!separate && gxtc::bits_of_triple(sz) > 64) || f.nbytes > 0x1ffffff0u)
            return gorder_gpu_run_xtc(h, x, atom_of_slot, first, last, stride, frame_index0, n_threads, batch_frames, nullptr);
        max_bytes = std::max(max_bytes, f.nbytes);
    }
    cudaSetDevice(h->device);
    if (!h->xtc_dev) h->xtc_dev = new GorderXtcDev();
    GorderXtcDev &D = *h->xtc_dev;
    const int na = h->s.n_atoms;
    const int B = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(batch_frames > 0 ? batch_frames : 32, h->max_batch), total));
    // 16-byte aligned streams with room for the reader's look-ahead; 1/8 of slack so that later, slightly larger frames of the
    // same trajectory do not force the pinned buffers to be re-allocated (frames that outgrow it do)
    size_t frame_cap = (((max_bytes + max_bytes / 8) + 4095) & ~(size_t)4095) + 16;
    if (h->xtc_dev && h->xtc_dev->frame_cap >= ((max_bytes + 15) & ~(size_t)15) + 16 && h->xtc_dev->n_traj_atoms == x->natoms) frame_cap = h->xtc_dev->frame_cap;
    const size_t marks_per_frame = (size_t)x->natoms / gxtc::kBookmarkEvery + 2;
    if (!D.magic_uploaded) { CK(cudaMemcpyToSymbol(gxtc::c_magic, gxtc::kMagic, sizeof(gxtc::kMagic))); D.magic_uploaded = true; }
    if (D.frames_cap < B || D.bytes_cap < (size_t)B * frame_cap || D.n_traj_atoms != x->natoms) {
        CK(cudaStreamSynchronize(h->stream));
        D.free_all();
        D = GorderXtcDev();
        D.magic_uploaded = true;
        D.frames_cap = B; D.bytes_cap = (size_t)B * frame_cap; D.frame_cap = frame_cap; D.n_traj_atoms = x->natoms; D.marks_per_frame = marks_per_frame;
        for (int i = 0; i < 2; i++) {
            CK(cudaMallocHost((void **)&D.h_bytes[i], D.bytes_cap)); CK(cudaMalloc((void **)&D.d_bytes[i], D.bytes_cap));
            CK(cudaMallocHost((void **)&D.h_frames[i], B * sizeof(gxtc::DevFrame))); CK(cudaMalloc((void **)&D.d_frames[i], B * sizeof(gxtc::DevFrame)));
            CK(cudaMallocHost((void **)&D.h_marks[i], B * marks_per_frame * sizeof(gxtc::Bookmark)));
            CK(cudaMalloc((void **)&D.d_marks[i], B * marks_per_frame * sizeof(gxtc::Bookmark)));
            CK(cudaMallocHost((void **)&D.h_box[i], B * 3 * sizeof(float))); CK(cudaMalloc((void **)&D.d_box[i], B * 3 * sizeof(float)));
            CK(cudaEventCreateWithFlags(&D.ev_copied[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&D.ev_done[i], cudaEventDisableTiming));
        }
        CK(cudaMalloc((void **)&D.d_slot_of_atom, (size_t)x->natoms * sizeof(int)));
        CK(cudaMalloc((void **)&D.d_xyz, (size_t)B * na * 3 * sizeof(float)));
        CK(cudaMemset(D.d_xyz, 0, (size_t)B * na * 3 * sizeof(float)));
    }
    {   // trajectory atom -> engine atom (atoms the analysis does not need are dropped by the decoder)
        std::vector<int> slot_of_atom((size_t)x->natoms, -1);
        for (int s = 0; s < na; s++) {
            const int a = atom_of_slot ? atom_of_slot[s] : s;
            if (a < 0 || a >= x->natoms) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "atom_of_slot outside the trajectory", a); return h->err_code; }
            if (h->slot_off[s] >= 0) slot_of_atom[(size_t)a] = s;
        }
        if (slot_of_atom != D.slot_cache) {   // unchanged between the calls of one run: keep the device copy
            CK(cudaStreamSynchronize(h->stream));
            CK(cudaMemcpy(D.d_slot_of_atom, slot_of_atom.data(), slot_of_atom.size() * sizeof(int), cudaMemcpyHostToDevice));
            D.slot_cache.swap(slot_of_atom);
        }
    }
    std::vector<int64_t> fi((size_t)B);
    long long moved = 0;
    std::atomic<int> bad{0}, unsupported{0};
    std::vector<int> n_groups((size_t)B);
    // Who walks the control bits?  The device (xtc_walk_kernel: the host only hands over the compressed bytes) unless a frame
    // could meet a small triple of more than 64 bits -- smallidx moves inside [first - 8, first + 8] of the frame header's
    // value (the writer's window) -- in which case the host walks, as in round 1, and hands such frames to the host decoder.
    bool dev_walk = !h->sw.xtc_host_walk;
    for (int64_t j = 0; dev_walk && j < total; j++)
        if (x->frames[(size_t)(first + j * stride)].smallidx + 8 > 64) dev_walk = false;
    auto stage = [&](int64_t j0, int nf, int buf) {   // compressed frames (+ bookmarks, when the host walks) -> pinned batch
        std::atomic<int> next{0};
        const int nt = std::max(1, std::min<int>(n_threads, nf));
        // a thread walks the control bits of up to kWalkWays frames in one loop (gxtc::Walker): the walk is a chain of dependent
        // steps, several chains overlap in the core (x2 per thread at 4 ways); not more ways than keep every thread busy
        const int ways = std::max(1, std::min(gxtc::kWalkWays, nf / nt));
        auto work = [&]() {
            std::vector<gxtc::Bookmark> marks[gxtc::kWalkWays];
            for (;;) {
                const int jb = next.fetch_add(ways);
                if (jb >= nf) break;
                const int m = std::min(ways, nf - jb);
                const gxtc::Frame *fr[gxtc::kWalkWays];
                std::vector<gxtc::Bookmark> *out[gxtc::kWalkWays];
                int bits[gxtc::kWalkWays], ngs[gxtc::kWalkWays];
                for (int k = 0; k < m; k++) {
                    const int j = jb + k;
                    const gxtc::Frame &f = x->frames[(size_t)(first + (j0 + j) * stride)];
                    if (f.box[1] != 0.0f || f.box[2] != 0.0f || f.box[3] != 0.0f || f.box[5] != 0.0f || f.box[6] != 0.0f || f.box[7] != 0.0f) bad = GORDER_ERR_NOT_ORTHOGONAL_BOX;
                    gxtc::DevFrame &d = D.h_frames[buf][j];
                    d.payload = (unsigned long long)j * frame_cap; d.bookmarks = (unsigned long long)j * marks_per_frame;
                    d.nbytes = (unsigned)f.nbytes; d.natoms = f.natoms;
                    unsigned sz[3];
                    for (int c = 0; c < 3; c++) { d.minint[c] = f.minint[c]; sz[c] = d.sizeint[c] = (unsigned)(f.maxint[c] - f.minint[c] + 1); }
                    if ((sz[0] | sz[1] | sz[2]) > 0xffffffu) { d.bitsize = 0; for (int c = 0; c < 3; c++) d.bitsint[c] = gxtc::bits_of(sz[c]); }
                    else { d.bitsize = gxtc::bits_of_triple(sz); d.bitsint[0] = d.bitsint[1] = d.bitsint[2] = 0; }
                    d.inv_precision = 1.0f / f.precision; d.smallidx = f.smallidx; d.pad_ = 0;
                    fr[k] = &f; bits[k] = d.bitsize ? d.bitsize : d.bitsint[0] + d.bitsint[1] + d.bitsint[2];
                    marks[k].clear(); out[k] = &marks[k];
                }
                if (!dev_walk) gxtc::bookmark_many(x->data, fr, bits, out, ngs, m);
                for (int k = 0; k < m; k++) {
                    const int j = jb + k, ng = dev_walk ? 0 : ngs[k];
                    const gxtc::Frame &f = *fr[k];
                    gxtc::DevFrame &d = D.h_frames[buf][j];
                    if (dev_walk) d.n_groups = 0;   // filled in by xtc_walk_kernel
                    else if (ng == -2) { unsupported = 1; d.n_groups = 0; }
                    else if (ng < 0 || marks[k].size() > marks_per_frame) { bad = GORDER_ERR_INVALID_ARGUMENT; d.n_groups = 0; }
                    else { d.n_groups = ng; memcpy(D.h_marks[buf] + d.bookmarks, marks[k].data(), marks[k].size() * sizeof(gxtc::Bookmark)); }
                    n_groups[(size_t)j] = d.n_groups;
                    unsigned char *dst = D.h_bytes[buf] + d.payload;
                    if (h->sw.xtc_nt_copy) gxtc::stream_copy(dst, x->data + f.payload, f.nbytes); else memcpy(dst, x->data + f.payload, f.nbytes);
                    // the readers look up to 8 bytes past a field: zeros behind the stream (the whole slack when it travels)
                    memset(dst + f.nbytes, 0, dev_walk ? std::min<size_t>(32, frame_cap - f.nbytes) : frame_cap - f.nbytes);
                    D.h_box[buf][3 * j] = f.box[0]; D.h_box[buf][3 * j + 1] = f.box[4]; D.h_box[buf][3 * j + 2] = f.box[8];
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    };
    int rc = GORDER_OK;
    int64_t j0 = 0;
    int nf = (int)std::min<int64_t>(B, total);
    CK(cudaEventSynchronize(D.ev_done[0]));
    stage(0, nf, 0);
    for (int buf = 0; j0 < total && !rc; buf ^= 1) {
        if (bad) { h->set_error(bad, bad == GORDER_ERR_NOT_ORTHOGONAL_BOX ? "simulation box is not orthogonal" : "corrupt XTC frame"); rc = h->err_code; break; }
        if (unsupported) {   // a frame of this batch needs > 64 bits per small triple: the host decoder takes over from here
            rc = gorder_gpu_run_xtc(h, x, atom_of_slot, first + j0 * stride, last, stride, frame_index0 + j0 * stride, n_threads, batch_frames, nullptr);
            break;
        }
        size_t nbytes = (size_t)nf * frame_cap, mbytes = (size_t)nf * marks_per_frame * sizeof(gxtc::Bookmark);
        if (dev_walk) {   // only the streams travel (+ 32 bytes of zeros each): no slack, no bookmarks
            nbytes = 0; mbytes = 0;
            for (int j = 0; j < nf; j++) {
                const size_t off = (size_t)D.h_frames[buf][j].payload, len = std::min(frame_cap, (((size_t)D.h_frames[buf][j].nbytes + 32) + 15) & ~(size_t)15);
                CK(cudaMemcpyAsync(D.d_bytes[buf] + off, D.h_bytes[buf] + off, len, cudaMemcpyHostToDevice, h->copy_stream));
                nbytes += len;
            }
        } else {
            CK(cudaMemcpyAsync(D.d_bytes[buf], D.h_bytes[buf], nbytes, cudaMemcpyHostToDevice, h->copy_stream));
            CK(cudaMemcpyAsync(D.d_marks[buf], D.h_marks[buf], mbytes, cudaMemcpyHostToDevice, h->copy_stream));
        }
        CK(cudaMemcpyAsync(D.d_frames[buf], D.h_frames[buf], nf * sizeof(gxtc::DevFrame), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaMemcpyAsync(D.d_box[buf], D.h_box[buf], nf * 3 * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(D.ev_copied[buf], h->copy_stream));
        moved += (long long)(nbytes + mbytes);
        CK(cudaStreamWaitEvent(h->stream, D.ev_copied[buf], 0));
        int max_groups = 1;
        if (dev_walk) {
            gxtc::xtc_walk_kernel<<<nf, 32, 0, h->stream>>>(D.d_bytes[buf], D.d_frames[buf], D.d_marks[buf], (unsigned)marks_per_frame, (unsigned)frame_cap, h->d_err, h->d_err_detail,
                                                          (long long)(first + j0 * stride), (long long)stride);
            h->n_launches++;
            max_groups = std::max(1, x->natoms / 4);   // the decoder strides: any grid is correct, this one covers ~5 atoms per group in one pass
        } else
            for (int j = 0; j < nf; j++) max_groups = std::max(max_groups, n_groups[(size_t)j]);
        dim3 grid((max_groups + 255) / 256, nf);
        gxtc::xtc_decode_kernel<<<grid, 256, 0, h->stream>>>(D.d_bytes[buf], D.d_frames[buf], D.d_marks[buf], D.d_slot_of_atom, na, D.d_xyz);
        h->n_launches++;
        CK(cudaGetLastError());
        for (int j = 0; j < nf; j++) fi[(size_t)j] = frame_index0 + (j0 + j) * stride;
        rc = gorder_gpu_submit_device(h, D.d_xyz, D.d_box[buf], fi.data(), nf);
        CK(cudaEventRecord(D.ev_done[buf], h->stream));
        // stage the next batch while the device works on this one
        const int64_t j1 = j0 + nf;
        const int nf1 = (int)std::min<int64_t>(B, total - j1);
        if (nf1 > 0 && !rc) {
            CK(cudaEventSynchronize(D.ev_done[buf ^ 1]));   // the other pinned buffer's previous batch has been consumed
            stage(j1, nf1, buf ^ 1);
        }
        j0 = j1; nf = nf1;
    }
    if (bytes_h2d) *bytes_h2d = moved;
    return rc;
}

}  // extern "C"
