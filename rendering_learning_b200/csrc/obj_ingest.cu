// OBJ ingest on the device: Wavefront OBJ text -> per-triangle arrays in HBM, and from there straight
// into the flat scene (vertex transform bake, TriVerts / TriShade packing, LBVH input boxes) without the triangles ever
// visiting the host.
//
// Replaces WavefrontObj::parse + to_object of both crates
//   RTC/src/io/wavefront_obj.rs:22-76 (parse), 78-187 (parse_vertex / parse_normal / parse_face / fan_triangulate)
//   OW/src/io/wavefront_obj.rs:32-104 (parse), 106-252 (parse_texture_coord, parse_face with vt)
// record for record: `v`, `vn`, `vt` (OW only), `f` (fan triangulation; v, v/t, v//n, v/t/n), `g`; anything else — and
// any record that fails to parse — only bumps `ignored`.
//
// Parallel formulation of a sequential parser:
//   1. k_mark_lines      one thread per byte: line starts (reader.lines(): split at '\n', a trailing '\r' dropped)
//   2. scan              exclusive prefix sums (hand-written three-phase scan, like the LBVH's radix sort: no CUB)
//   3. k_parse_lines     one thread per line: split_once(' '), classify the head, parse the numbers of v / vn / vt with an
//                        EXACT decimal -> f64 conversion (up to 2^53 x 10^+-22: one IEEE multiply or divide of two exactly
//                        representable numbers = correctly rounded; otherwise big-integer division with a sticky bit;
//                        both give what `str::parse::<f64>` returns.  Beyond 19 significant digits or |exp10| > 60 the
//                        whole ingest fails with RL_E_UNSUPPORTED instead of rounding differently), count the triangles
//                        of an f record
//   4. scans over lines  positions of every v / vn / vt / triangle in their output arrays = the order the sequential
//                        parser would have pushed them in
//   5. k_emit            v / vn / vt records to their arrays; f records re-tokenised, indices resolved against the records
//                        read BEFORE that line (an index beyond them is the reference's out-of-bounds panic -> error)
//   6. groups            `obj.groups.insert(name, tris)` REPLACES the triangles of an earlier group with the same name, and
//                        to_object() walks the map: the (few) `g` records go to the host, which decides which segments
//                        survive and in which order, and one gather kernel puts the triangles in that order.
#include <cstring>
#include <string>
#include <vector>

#include "obj_ingest.h"

namespace rl {
namespace {

enum : int { LK_IGNORED = 0, LK_V = 1, LK_VN = 2, LK_VT = 3, LK_F = 4, LK_G = 5 };
constexpr int ERR_PRECISION = 1, ERR_INDEX = 2;

__device__ __forceinline__ bool is_ws(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// ---- scan ------------------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_PER = 4, SCAN_TILE = SCAN_T * SCAN_PER;

__global__ void k_scan_tile(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tile_sums, long long n) {
    __shared__ int sm[SCAN_T];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER;
    int v[SCAN_PER], sum = 0;
    for (int k = 0; k < SCAN_PER; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        sum += v[k];
    }
    sm[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < SCAN_T; off <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
        int t = threadIdx.x >= off ? sm[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += t;
        __syncthreads();
    }
    int run = sm[threadIdx.x] - sum;  // exclusive prefix of this thread within the tile
    for (int k = 0; k < SCAN_PER; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_T - 1 && tile_sums) tile_sums[blockIdx.x] = sm[SCAN_T - 1];
}
__global__ void k_scan_add(int* __restrict__ out, const int* __restrict__ tile_prefix, long long n) {
    const long long i = (long long)blockIdx.x * SCAN_TILE + threadIdx.x;
    const int add = tile_prefix[blockIdx.x];
    for (int k = 0; k < SCAN_PER; k++) {
        long long j = i + (long long)k * SCAN_T;
        if (j < n) out[j] += add;
    }
}

// exclusive scan of n ints (in -> out, may alias); `scratch` holds ceil(n / TILE) + ceil(that / TILE) + ... ints
cudaError_t scan_exclusive(const int* in, int* out, long long n, int* scratch, cudaStream_t s, int* launches) {
    if (n <= 0) return cudaSuccess;
    const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tile<<<(unsigned)tiles, SCAN_T, 0, s>>>(in, out, tiles > 1 ? scratch : nullptr, n);
    if (launches) ++*launches;
    if (tiles > 1) {
        cudaError_t e = scan_exclusive(scratch, scratch, tiles, scratch + tiles, s, launches);
        if (e != cudaSuccess) return e;
        k_scan_add<<<(unsigned)tiles, SCAN_T, 0, s>>>(out, scratch, n);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}
long long scan_scratch_ints(long long n) {
    long long total = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += n;
    }
    return total + 1;
}

// ---- 1. lines --------------------------------------------------------------------------------------------------------
__global__ void k_mark_lines(const char* __restrict__ text, long long n, int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || text[i - 1] == '\n') ? 1 : 0;
}
__global__ void k_line_starts(const int* __restrict__ flag, const int* __restrict__ pos, long long n, long long* __restrict__ start) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) start[pos[i]] = i;
}

// ---- number parsing ----------------------------------------------------------------------------------------------------
// Exact decimal -> f64 for the numbers the one-operation fast path cannot take (|exp10| > 22 or more than 53 bits of
// mantissa): mant * 10^exp10 as a ratio of two big integers, 55-56 quotient bits by restoring division, the remainder
// as the sticky bit, round to nearest even.  12 x 32-bit words cover |exp10| <= 60 with a 60-bit mantissa.  Slow (a few
// thousand instructions) and rare: an OBJ file carries a handful of such numbers, if any.
constexpr int BIGW = 12;
struct Big {
    unsigned w[BIGW];  // little endian
};
__device__ int big_bitlen(const Big& a) {
    for (int i = BIGW - 1; i >= 0; i--)
        if (a.w[i]) return 32 * i + (32 - __clz(a.w[i]));
    return 0;
}
__device__ void big_mul_small(Big& a, unsigned m) {
    unsigned long long carry = 0;
    for (int i = 0; i < BIGW; i++) {
        unsigned long long t = (unsigned long long)a.w[i] * m + carry;
        a.w[i] = (unsigned)t;
        carry = t >> 32;
    }
}
__device__ void big_shl(Big& a, int s) {
    const int ws = s / 32, bs = s % 32;
    for (int i = BIGW - 1; i >= 0; i--) {
        unsigned lo = i - ws >= 0 ? a.w[i - ws] : 0u, lo2 = (bs && i - ws - 1 >= 0) ? a.w[i - ws - 1] : 0u;
        a.w[i] = bs ? ((lo << bs) | (lo2 >> (32 - bs))) : lo;
    }
}
__device__ bool big_geq(const Big& a, const Big& b) {
    for (int i = BIGW - 1; i >= 0; i--)
        if (a.w[i] != b.w[i]) return a.w[i] > b.w[i];
    return true;
}
__device__ void big_sub(Big& a, const Big& b) {
    long long borrow = 0;
    for (int i = 0; i < BIGW; i++) {
        long long t = (long long)a.w[i] - b.w[i] - borrow;
        borrow = t < 0;
        a.w[i] = (unsigned)t;
    }
}
__device__ double decimal_to_double_exact(unsigned long long mant, int exp10) {
    Big num{}, den{};
    num.w[0] = (unsigned)mant;
    num.w[1] = (unsigned)(mant >> 32);
    den.w[0] = 1u;
    for (int k = 0; k < (exp10 < 0 ? -exp10 : exp10); k++) big_mul_small(exp10 < 0 ? den : num, 10u);
    // scale so that floor(num / den) has 55 or 56 bits
    const int s = 55 - (big_bitlen(num) - big_bitlen(den));
    if (s >= 0) big_shl(num, s); else big_shl(den, -s);
    // restoring division, most significant bit first; the quotient fits 64 bits
    Big rem{};
    unsigned long long q = 0;
    for (int bit = big_bitlen(num) - 1; bit >= 0; bit--) {
        big_shl(rem, 1);
        rem.w[0] |= (num.w[bit / 32] >> (bit % 32)) & 1u;
        q <<= 1;
        if (big_geq(rem, den)) {
            big_sub(rem, den);
            q |= 1ull;
        }
    }
    const bool sticky = big_bitlen(rem) != 0;
    const int qbits = 64 - __clzll((long long)q);
    const int shift = qbits - 53;  // 2 or 3
    unsigned long long m53 = q >> shift;
    const unsigned long long low = q & ((1ull << shift) - 1ull), half = 1ull << (shift - 1);
    if (low > half || (low == half && (sticky || (m53 & 1ull)))) m53++;
    return ldexp((double)m53, shift - s);  // m53 <= 2^53: exact; the scaling by a power of two is exact in the normal range
}

// `str::parse::<f64>` on [b, e): true + value, or false.  *inexact is set when the number is valid but outside the
// exactly-convertible range (see the file header).
__device__ bool parse_f64(const char* __restrict__ t, long long b, long long e, double* out, int* inexact) {
    if (b >= e) return false;
    bool neg = false;
    if (t[b] == '+' || t[b] == '-') {
        neg = t[b] == '-';
        b++;
    }
    if (b >= e) return false;
    // inf / infinity / nan, case-insensitive
    auto lower = [](char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; };
    const long long len = e - b;
    if (len == 3 || len == 8) {
        const char* w = len == 3 ? "inf" : "infinity";
        bool same = true;
        for (long long k = 0; k < len; k++) same = same && lower(t[b + k]) == w[k];
        if (same) {
            *out = neg ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL);
            return true;
        }
        if (len == 3 && lower(t[b]) == 'n' && lower(t[b + 1]) == 'a' && lower(t[b + 2]) == 'n') {
            *out = __longlong_as_double(0x7ff8000000000000LL);
            return true;
        }
    }
    unsigned long long mant = 0;
    int digits = 0, sig = 0, exp10 = 0;
    bool seen_dot = false, too_long = false;
    long long i = b;
    for (; i < e; i++) {
        const char c = t[i];
        if (c >= '0' && c <= '9') {
            digits++;
            if (mant != 0 || c != '0') {
                if (sig < 19) {  // 10^19 < 2^64
                    mant = mant * 10ull + (unsigned long long)(c - '0');
                    sig++;
                    if (seen_dot) exp10--;
                } else {  // beyond what fits: only harmless if every further digit is a zero
                    if (c != '0') too_long = true;
                    if (!seen_dot) exp10++;
                }
            } else if (seen_dot) {
                exp10--;
            }
        } else if (c == '.' && !seen_dot) {
            seen_dot = true;
        } else {
            break;
        }
    }
    if (digits == 0) return false;
    if (i < e) {
        if (t[i] != 'e' && t[i] != 'E') return false;
        i++;
        bool eneg = false;
        if (i < e && (t[i] == '+' || t[i] == '-')) {
            eneg = t[i] == '-';
            i++;
        }
        if (i >= e) return false;
        int ev = 0;
        for (; i < e; i++) {
            if (t[i] < '0' || t[i] > '9') return false;
            if (ev < 100000) ev = ev * 10 + (t[i] - '0');
        }
        exp10 += eneg ? -ev : ev;
    }
    double v;
    if (mant == 0) {
        v = 0.0;
    } else {
        // trailing zeros of the mantissa move into the exponent (1.500000 -> 15 x 10^-1), which keeps more numbers exact
        while (mant % 10ull == 0ull && exp10 < 0) { mant /= 10ull; exp10++; }
        if (too_long || exp10 < -60 || exp10 > 60) {
            *inexact = 1;
            return true;
        }
        if (mant <= (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
            // fast path (Clinger): both operands are exactly representable, ONE IEEE operation = correctly rounded
            const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
            const double m = (double)mant;
            v = exp10 < 0 ? __ddiv_rn(m, p10[-exp10]) : __dmul_rn(m, p10[exp10]);
        } else {
            v = decimal_to_double_exact(mant, exp10);  // e.g. spot_triangulated.obj's "-4.33681e-19"
        }
    }
    *out = neg ? -v : v;
    return true;
}

// `str::parse::<usize>`: optional '+', then digits only
__device__ bool parse_usize(const char* __restrict__ t, long long b, long long e, long long* out) {
    if (b < e && t[b] == '+') b++;
    if (b >= e) return false;
    long long v = 0;
    for (long long i = b; i < e; i++) {
        if (t[i] < '0' || t[i] > '9') return false;
        if (v > (1ll << 40)) return false;
        v = v * 10 + (t[i] - '0');
    }
    *out = v;
    return true;
}

struct Token {
    long long b, e;
};
// next whitespace-separated token of [*pos, end)
__device__ __forceinline__ bool next_token(const char* __restrict__ t, long long* pos, long long end, Token* tok) {
    long long i = *pos;
    while (i < end && is_ws(t[i])) i++;
    if (i >= end) return false;
    tok->b = i;
    while (i < end && !is_ws(t[i])) i++;
    tok->e = i;
    *pos = i;
    return true;
}

// one face token "v", "v/t", "v//n", "v/t/n" -> indices (0 = absent); false = the whole face record is ignored
__device__ bool parse_face_token(const char* __restrict__ t, Token tok, int flavor, long long* v, long long* vt, long long* vn) {
    long long part_b[3], part_e[3];
    int parts = 0;
    long long s = tok.b;
    for (long long i = tok.b; i <= tok.e; i++) {
        if (i == tok.e || t[i] == '/') {
            if (parts == 3) return false;  // more than three '/'-separated parts
            part_b[parts] = s;
            part_e[parts] = i;
            parts++;
            s = i + 1;
        }
    }
    *v = *vt = *vn = 0;
    if (!parse_usize(t, part_b[0], part_e[0], v)) return false;
    if (flavor == RL_FLAVOR_RTC) {  // RTC/src/io/wavefront_obj.rs:116-137: vt is never read; an unparsable vn voids the face
        if (parts == 3 && !parse_usize(t, part_b[2], part_e[2], vn)) return false;
    } else {  // OW/src/io/wavefront_obj.rs:159-183: unparsable vt / vn are simply absent
        if (parts >= 2 && part_b[1] < part_e[1] && !parse_usize(t, part_b[1], part_e[1], vt)) *vt = 0;
        if (parts == 3 && !parse_usize(t, part_b[2], part_e[2], vn)) *vn = 0;
    }
    return true;
}

// ---- 3. per-line parse -------------------------------------------------------------------------------------------------
struct LineOut {
    int* kind;        // LK_*
    int* is_v;        // 1 for an accepted v record (scanned into positions), same for vn / vt / g / ignored
    int* is_vn;
    int* is_vt;
    int* is_g;
    int* n_tris;      // triangles an accepted f record produces
    double* vals;     // [line][3] numbers of an accepted v / vn / vt record
};

__global__ void k_parse_lines(const char* __restrict__ text, long long n, const long long* __restrict__ start, int n_lines,
                              int flavor, LineOut o, int* __restrict__ err) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines) return;
    long long b = start[li], e = li + 1 < n_lines ? start[li + 1] : n;
    if (e > b && text[e - 1] == '\n') e--;
    if (e > b && text[e - 1] == '\r') e--;
    int kind = LK_IGNORED, tris = 0;
    double v3[3] = {0.0, 0.0, 0.0};
    long long sp = b;
    while (sp < e && text[sp] != ' ') sp++;
    if (sp < e) {  // split_once(' ')
        const long long hl = sp - b;
        long long tb = sp + 1, te = e;
        while (tb < te && is_ws(text[tb])) tb++;  // tail.trim()
        while (te > tb && is_ws(text[te - 1])) te--;
        const char h0 = text[b], h1 = hl > 1 ? text[b + 1] : 0;
        int head = LK_IGNORED;
        if (hl == 1 && h0 == 'v') head = LK_V;
        else if (hl == 1 && h0 == 'f') head = LK_F;
        else if (hl == 1 && h0 == 'g') head = LK_G;
        else if (hl == 2 && h0 == 'v' && h1 == 'n') head = LK_VN;
        else if (hl == 2 && h0 == 'v' && h1 == 't' && flavor == RL_FLAVOR_OW) head = LK_VT;
        if (head == LK_G) {
            kind = LK_G;
        } else if (head == LK_V || head == LK_VN || head == LK_VT) {
            long long pos = tb;
            Token tok;
            int cnt = 0, inexact = 0;
            bool ok = true;
            while (next_token(text, &pos, te, &tok)) {
                double x;
                if (!parse_f64(text, tok.b, tok.e, &x, &inexact)) { ok = false; break; }
                if (cnt < 3) v3[cnt] = x;
                cnt++;
            }
            if (head == LK_VT) {  // parse_texture_coord: 0 numbers -> None, 1 -> (x, 0.0), more -> the first two
                ok = ok && cnt >= 1;
                if (cnt == 1) v3[1] = 0.0;
            } else {
                ok = ok && cnt == 3;
            }
            if (ok) {
                kind = head;
                if (inexact) atomicOr(err, ERR_PRECISION);
            }
        } else if (head == LK_F) {
            long long pos = tb;
            Token tok;
            int cnt = 0;
            bool ok = true;
            while (next_token(text, &pos, te, &tok)) {
                long long a, c, d;
                if (!parse_face_token(text, tok, flavor, &a, &c, &d)) { ok = false; break; }
                cnt++;
            }
            if (ok && cnt >= 3) {
                kind = LK_F;
                tris = cnt - 2;
            }
        }
    }
    o.kind[li] = kind;
    o.is_v[li] = kind == LK_V;
    o.is_vn[li] = kind == LK_VN;
    o.is_vt[li] = kind == LK_VT;
    o.is_g[li] = kind == LK_G;
    o.n_tris[li] = tris;
    o.vals[3 * (size_t)li + 0] = v3[0];
    o.vals[3 * (size_t)li + 1] = v3[1];
    o.vals[3 * (size_t)li + 2] = v3[2];
}

// ---- 5. emit -------------------------------------------------------------------------------------------------------------
__global__ void k_emit_records(int n_lines, const int* __restrict__ kind, const double* __restrict__ vals,
                               const int* __restrict__ pv, const int* __restrict__ pn, const int* __restrict__ pt,
                               double* __restrict__ verts, double* __restrict__ norms, double* __restrict__ tex) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines) return;
    const double* v = vals + 3 * (size_t)li;
    if (kind[li] == LK_V) {
        double* d = verts + 3 * (size_t)pv[li];
        d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
    } else if (kind[li] == LK_VN) {
        double* d = norms + 3 * (size_t)pn[li];
        d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
    } else if (kind[li] == LK_VT) {
        double* d = tex + 2 * (size_t)pt[li];
        d[0] = v[0]; d[1] = v[1];
    }
}

__global__ void k_emit_faces(const char* __restrict__ text, long long n, const long long* __restrict__ start, int n_lines,
                             int flavor, const int* __restrict__ kind, const int* __restrict__ pv, const int* __restrict__ pn,
                             const int* __restrict__ pt, const int* __restrict__ ptri, const double* __restrict__ verts,
                             const double* __restrict__ norms, const double* __restrict__ tex, ObjMesh m, int* __restrict__ err) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines || kind[li] != LK_F) return;
    long long b = start[li], e = li + 1 < n_lines ? start[li + 1] : n;
    if (e > b && text[e - 1] == '\n') e--;
    if (e > b && text[e - 1] == '\r') e--;
    long long pos = b + 2;  // past "f "
    Token tok;
    // records read BEFORE this line: what the sequential parser's vectors hold when it reaches it
    const long long have_v = pv[li], have_n = pn[li], have_t = pt[li];
    long long v0 = 0, t0 = 0, n0 = 0, vp = 0, tp = 0, np = 0;
    int cnt = 0, tri = ptri[li];
    while (next_token(text, &pos, e, &tok)) {
        long long v, vt, vn;
        parse_face_token(text, tok, flavor, &v, &vt, &vn);  // validated by k_parse_lines
        // read_vertices[vi - 1] etc.: index 0 underflows, an index past the records read so far is out of bounds — both
        // panic in the reference; here the ingest fails as a whole
        if (v < 1 || v > have_v || vt > have_t || vn > have_n) {
            atomicOr(err, ERR_INDEX);
            return;
        }
        if (cnt == 0) { v0 = v; t0 = vt; n0 = vn; }
        if (cnt >= 2) {  // fan: (first, previous, current)
            const long long vi[3] = {v0, vp, v}, ti[3] = {t0, tp, vt}, ni[3] = {n0, np, vn};
            const bool has_n = ni[0] && ni[1] && ni[2], has_t = ti[0] && ti[1] && ti[2];
            for (int k = 0; k < 3; k++) {
                for (int c = 0; c < 3; c++) {
                    m.tri_p[9 * (size_t)tri + 3 * k + c] = verts[3 * (size_t)(vi[k] - 1) + c];
                    m.tri_n[9 * (size_t)tri + 3 * k + c] = has_n ? norms[3 * (size_t)(ni[k] - 1) + c] : 0.0;
                }
                for (int c = 0; c < 2; c++) m.tri_uv[6 * (size_t)tri + 2 * k + c] = has_t ? tex[2 * (size_t)(ti[k] - 1) + c] : 0.0;
            }
            m.tri_flags[tri] = (unsigned char)((has_n ? 1 : 0) | (has_t ? 2 : 0));
            tri++;
        }
        vp = v; tp = vt; np = vn;
        cnt++;
    }
}

// ---- 6. group order ------------------------------------------------------------------------------------------------------
struct Segment {
    int src, count, dst;
};
__global__ void k_gather_segments(const Segment* __restrict__ segs, int n_segs, ObjMesh src, ObjMesh dst) {
    const int sgi = blockIdx.y;
    if (sgi >= n_segs) return;
    const Segment sg = segs[sgi];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.count; i += gridDim.x * blockDim.x) {
        const size_t a = (size_t)sg.src + i, b = (size_t)sg.dst + i;
        for (int c = 0; c < 9; c++) {
            dst.tri_p[9 * b + c] = src.tri_p[9 * a + c];
            dst.tri_n[9 * b + c] = src.tri_n[9 * a + c];
        }
        for (int c = 0; c < 6; c++) dst.tri_uv[6 * b + c] = src.tri_uv[6 * a + c];
        dst.tri_flags[b] = src.tri_flags[a];
    }
}

// ---- mesh instance -> flat scene -----------------------------------------------------------------------------------------
// exactly the arithmetic of flatten.cpp's aff_point / aff_normal / push_triangle, in f64 with the roundings pinned
// (no FMA contraction), so a device-ingested mesh yields the SAME f32 TriVerts / TriShade / boxes as the host path
__device__ __forceinline__ double dot3_rn(const double* r, const double* p) {
    return __dadd_rn(__dadd_rn(__dmul_rn(r[0], p[0]), __dmul_rn(r[1], p[1])), __dmul_rn(r[2], p[2]));
}
__global__ void k_mesh_instance(ObjMesh m, int n, MeshInstance inst, TriVerts* __restrict__ tv, TriShade* __restrict__ ts,
                                float* __restrict__ aabb, int* __restrict__ refs, int* __restrict__ node_ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double P[3][3], N[3][3];
    const unsigned char fl = m.tri_flags[i];
    const bool has_n = fl & 1, has_uv = (fl & 2) && inst.flavor == RL_FLAVOR_OW;
    for (int k = 0; k < 3; k++) {
        const double* p = m.tri_p + 9 * (size_t)i + 3 * k;
        for (int r = 0; r < 3; r++) P[k][r] = __dadd_rn(dot3_rn(inst.fwd[r], p), inst.fwd[r][3]);  // aff_point
    }
    if (has_n) {
        for (int k = 0; k < 3; k++) {  // aff_normal: inv^T * n, NOT normalised (the kernels normalise the interpolated normal)
            const double* q = m.tri_n + 9 * (size_t)i + 3 * k;
            for (int c = 0; c < 3; c++)
                N[k][c] = __dadd_rn(__dadd_rn(__dmul_rn(inst.inv[0][c], q[0]), __dmul_rn(inst.inv[1][c], q[1])), __dmul_rn(inst.inv[2][c], q[2]));
        }
    } else if (inst.flavor == RL_FLAVOR_RTC) {
        // Triangle::flat (triangle.rs:30-42): normalize(e2 x e1) in OBJECT space, then inv^T and normalised again
        const double* q = m.tri_p + 9 * (size_t)i;
        double e1[3], e2[3], c[3], w[3];
        for (int k = 0; k < 3; k++) { e1[k] = __dsub_rn(q[3 + k], q[k]); e2[k] = __dsub_rn(q[6 + k], q[k]); }
        c[0] = __dsub_rn(__dmul_rn(e2[1], e1[2]), __dmul_rn(e2[2], e1[1]));
        c[1] = __dsub_rn(__dmul_rn(e2[2], e1[0]), __dmul_rn(e2[0], e1[2]));
        c[2] = __dsub_rn(__dmul_rn(e2[0], e1[1]), __dmul_rn(e2[1], e1[0]));
        double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(c[0], c[0]), __dmul_rn(c[1], c[1])), __dmul_rn(c[2], c[2])));
        for (int k = 0; k < 3; k++) c[k] = __ddiv_rn(c[k], len);
        for (int k = 0; k < 3; k++)
            w[k] = __dadd_rn(__dadd_rn(__dmul_rn(inst.inv[0][k], c[0]), __dmul_rn(inst.inv[1][k], c[1])), __dmul_rn(inst.inv[2][k], c[2]));
        len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(w[0], w[0]), __dmul_rn(w[1], w[1])), __dmul_rn(w[2], w[2])));
        for (int k = 0; k < 3; k++) N[0][k] = N[1][k] = N[2][k] = __ddiv_rn(w[k], len);
    } else {
        // Plane::new (flat/plane.rs:23-28): n = normalize(u x v) of the TRANSFORMED vertices
        double u[3], v[3], c[3];
        for (int k = 0; k < 3; k++) { u[k] = __dsub_rn(P[1][k], P[0][k]); v[k] = __dsub_rn(P[2][k], P[0][k]); }
        c[0] = __dsub_rn(__dmul_rn(u[1], v[2]), __dmul_rn(u[2], v[1]));
        c[1] = __dsub_rn(__dmul_rn(u[2], v[0]), __dmul_rn(u[0], v[2]));
        c[2] = __dsub_rn(__dmul_rn(u[0], v[1]), __dmul_rn(u[1], v[0]));
        const double len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(c[0], c[0]), __dmul_rn(c[1], c[1])), __dmul_rn(c[2], c[2])));
        for (int k = 0; k < 3; k++) N[0][k] = N[1][k] = N[2][k] = __ddiv_rn(c[k], len);
    }
    const int idx = inst.tri_first + i;
    const int flags = (has_n ? 1 : 0) | (has_uv ? 2 : 0) | (inst.xf << 8);
    TriVerts v;
    v.p0 = make_float4((float)P[0][0], (float)P[0][1], (float)P[0][2], __int_as_float(inst.material));
    v.p1 = make_float4((float)P[1][0], (float)P[1][1], (float)P[1][2], __int_as_float(inst.node));
    v.p2 = make_float4((float)P[2][0], (float)P[2][1], (float)P[2][2], __int_as_float(flags));
    tv[idx] = v;
    const double* uv = m.tri_uv + 6 * (size_t)i;
    TriShade s;
    s.s0 = make_float4((float)N[0][0], (float)N[0][1], (float)N[0][2], has_uv ? (float)uv[0] : 0.0f);
    s.s1 = make_float4((float)N[1][0], (float)N[1][1], (float)N[1][2], has_uv ? (float)uv[1] : 0.0f);
    s.s2 = make_float4((float)N[2][0], (float)N[2][1], (float)N[2][2], has_uv ? (float)uv[2] : 0.0f);
    s.s3 = make_float4(has_uv ? (float)uv[3] : 0.0f, has_uv ? (float)uv[4] : 0.0f, has_uv ? (float)uv[5] : 0.0f, 0.0f);
    ts[idx] = s;
    // LBVH input: the box of the f32-rounded vertices the device will actually intersect, padded by 4 ulp exactly like
    // flatten.cpp push_aabb (thin / touching geometry stays inside its box in the f32 slab test)
    const int bi = inst.bvh_first + i;
    for (int k = 0; k < 3; k++) {
        const float a = (&v.p0.x)[k], b = (&v.p1.x)[k], c = (&v.p2.x)[k];
        const float lo = fminf(a, fminf(b, c)), hi = fmaxf(a, fmaxf(b, c));
        aabb[6 * (size_t)bi + k] = __fsub_rn(lo, __fmul_rn(__fmul_rn(4.0f, 1.1920929e-7f), fmaxf(fabsf(lo), 1e-3f)));
        aabb[6 * (size_t)bi + 3 + k] = __fadd_rn(hi, __fmul_rn(__fmul_rn(4.0f, 1.1920929e-7f), fmaxf(fabsf(hi), 1e-3f)));
    }
    refs[bi] = make_ref(REF_TRI, idx);
    node_ids[bi] = inst.report_node;
}

// object-space bounds of the parsed points (the flattener needs the mesh's extent without seeing its triangles)
__global__ void k_mesh_bounds(const double* __restrict__ tri_p, int n_points, double* __restrict__ out6) {
    __shared__ double lo[3][256], hi[3][256];
    double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
    for (int i = threadIdx.x; i < n_points; i += blockDim.x)
        for (int c = 0; c < 3; c++) {
            const double v = tri_p[3 * (size_t)i + c];
            l[c] = fmin(l[c], v);
            h[c] = fmax(h[c], v);
        }
    for (int c = 0; c < 3; c++) { lo[c][threadIdx.x] = l[c]; hi[c][threadIdx.x] = h[c]; }
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off)
            for (int c = 0; c < 3; c++) {
                lo[c][threadIdx.x] = fmin(lo[c][threadIdx.x], lo[c][threadIdx.x + off]);
                hi[c][threadIdx.x] = fmax(hi[c][threadIdx.x], hi[c][threadIdx.x + off]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 3) { out6[threadIdx.x] = lo[threadIdx.x][0]; out6[3 + threadIdx.x] = hi[threadIdx.x][0]; }
}

template <class T>
cudaError_t dev_alloc(T** p, size_t count) {
    return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
}

}  // namespace

void obj_mesh_free(ObjMesh* m) {
    if (m->tri_p) cudaFree(m->tri_p);
    if (m->tri_n) cudaFree(m->tri_n);
    if (m->tri_uv) cudaFree(m->tri_uv);
    if (m->tri_flags) cudaFree(m->tri_flags);
    *m = ObjMesh{};
}

cudaError_t launch_mesh_instance(const ObjMesh& m, const MeshInstance& inst, TriVerts* tv, TriShade* ts, float* aabb, int* refs,
                                 int* node_ids, cudaStream_t s) {
    if (m.n_triangles <= 0) return cudaSuccess;
    k_mesh_instance<<<(m.n_triangles + 127) / 128, 128, 0, s>>>(m, m.n_triangles, inst, tv, ts, aabb, refs, node_ids);
    return cudaGetLastError();
}

#define OBJ_CK(x)                                    \
    do {                                             \
        cudaError_t e_ = (x);                        \
        if (e_ != cudaSuccess) {                     \
            *err = std::string(#x) + ": " + cudaGetErrorString(e_); \
            rc = RL_E_CUDA;                          \
            goto done;                               \
        }                                            \
    } while (0)

int obj_parse_device(const char* text, uint64_t len, int flavor, cudaStream_t s, ObjMesh* out, rl_obj_info* info,
                     std::string* err) {
    int rc = RL_OK;
    *out = ObjMesh{};
    *info = rl_obj_info{};
    char* d_text = nullptr;
    int *flag = nullptr, *pos = nullptr, *scratch = nullptr, *derr = nullptr;
    long long* start = nullptr;
    int *kind = nullptr, *is_v = nullptr, *is_vn = nullptr, *is_vt = nullptr, *is_g = nullptr, *n_tris = nullptr;
    double *vals = nullptr, *verts = nullptr, *norms = nullptr, *tex = nullptr;
    Segment* d_segs = nullptr;
    ObjMesh raw{};
    const long long n = (long long)len;
    int launches = 0;
    if (len > (1ull << 31) - 4096) {
        *err = "OBJ text larger than 2 GiB";
        return RL_E_UNSUPPORTED;
    }
    if (n == 0) {
        info->n_groups = 1;
        return RL_OK;
    }
    {
        OBJ_CK(dev_alloc(&d_text, (size_t)n));
        OBJ_CK(cudaMemcpyAsync(d_text, text, (size_t)n, cudaMemcpyHostToDevice, s));
        OBJ_CK(dev_alloc(&flag, (size_t)n));
        OBJ_CK(dev_alloc(&pos, (size_t)n));
        OBJ_CK(dev_alloc(&scratch, (size_t)scan_scratch_ints(n)));
        OBJ_CK(dev_alloc(&derr, 1));
        OBJ_CK(cudaMemsetAsync(derr, 0, sizeof(int), s));
        k_mark_lines<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_text, n, flag);
        launches++;
        OBJ_CK(scan_exclusive(flag, pos, n, scratch, s, &launches));
        int last_flag = 0, last_pos = 0;
        OBJ_CK(cudaMemcpyAsync(&last_flag, flag + n - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        OBJ_CK(cudaMemcpyAsync(&last_pos, pos + n - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        OBJ_CK(cudaStreamSynchronize(s));
        const int n_lines = last_pos + last_flag;
        OBJ_CK(dev_alloc(&start, (size_t)n_lines));
        k_line_starts<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flag, pos, n, start);
        launches++;
        OBJ_CK(dev_alloc(&kind, (size_t)n_lines));
        OBJ_CK(dev_alloc(&is_v, (size_t)n_lines));
        OBJ_CK(dev_alloc(&is_vn, (size_t)n_lines));
        OBJ_CK(dev_alloc(&is_vt, (size_t)n_lines));
        OBJ_CK(dev_alloc(&is_g, (size_t)n_lines));
        OBJ_CK(dev_alloc(&n_tris, (size_t)n_lines));
        OBJ_CK(dev_alloc(&vals, 3 * (size_t)n_lines));
        LineOut lo{kind, is_v, is_vn, is_vt, is_g, n_tris, vals};
        const unsigned lb = (unsigned)((n_lines + 127) / 128);
        k_parse_lines<<<lb, 128, 0, s>>>(d_text, n, start, n_lines, flavor, lo, derr);
        launches++;
        // positions = exclusive scans over the lines (in place); totals = last prefix + last value
        int tot[5], lastv[5];
        int* arrs[5] = {is_v, is_vn, is_vt, is_g, n_tris};
        for (int k = 0; k < 5; k++) OBJ_CK(cudaMemcpyAsync(&lastv[k], arrs[k] + n_lines - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        for (int k = 0; k < 5; k++) OBJ_CK(scan_exclusive(arrs[k], arrs[k], n_lines, scratch, s, &launches));
        for (int k = 0; k < 5; k++) OBJ_CK(cudaMemcpyAsync(&tot[k], arrs[k] + n_lines - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        std::vector<int> h_kind((size_t)n_lines), h_g((size_t)n_lines), h_tri((size_t)n_lines);
        std::vector<long long> h_start((size_t)n_lines);
        OBJ_CK(cudaMemcpyAsync(h_kind.data(), kind, sizeof(int) * (size_t)n_lines, cudaMemcpyDeviceToHost, s));
        OBJ_CK(cudaStreamSynchronize(s));
        for (int k = 0; k < 5; k++) tot[k] += lastv[k];
        const int nv = tot[0], nn = tot[1], nt = tot[2], ng = tot[3], ntri = tot[4];
        int ignored = 0;
        for (int k : h_kind) ignored += k == LK_IGNORED;
        info->n_vertices = nv;
        info->n_normals = nn;
        info->n_texcoords = nt;
        info->n_groups = ng + 1;
        info->ignored = ignored;
        OBJ_CK(dev_alloc(&verts, 3 * (size_t)nv));
        OBJ_CK(dev_alloc(&norms, 3 * (size_t)nn));
        OBJ_CK(dev_alloc(&tex, 2 * (size_t)nt));
        raw.n_triangles = ntri;
        OBJ_CK(dev_alloc(&raw.tri_p, 9 * (size_t)ntri));
        OBJ_CK(dev_alloc(&raw.tri_n, 9 * (size_t)ntri));
        OBJ_CK(dev_alloc(&raw.tri_uv, 6 * (size_t)ntri));
        OBJ_CK(dev_alloc(&raw.tri_flags, (size_t)ntri));
        k_emit_records<<<lb, 128, 0, s>>>(n_lines, kind, vals, is_v, is_vn, is_vt, verts, norms, tex);
        k_emit_faces<<<lb, 128, 0, s>>>(d_text, n, start, n_lines, flavor, kind, is_v, is_vn, is_vt, n_tris, verts, norms, tex, raw, derr);
        launches += 2;
        int h_err = 0;
        OBJ_CK(cudaMemcpyAsync(&h_err, derr, sizeof(int), cudaMemcpyDeviceToHost, s));
        // ---- groups: which segments survive `groups.insert`, and in which order to_object walks them ----
        // (the reference iterates a HashMap, i.e. in an unspecified order; the host mirror's order is the map's insertion
        //  order, a re-inserted name keeping its first position: that order is used here too)
        std::vector<Segment> segs;
        if (ng > 0) {
            OBJ_CK(cudaMemcpyAsync(h_tri.data(), n_tris, sizeof(int) * (size_t)n_lines, cudaMemcpyDeviceToHost, s));
            OBJ_CK(cudaMemcpyAsync(h_start.data(), start, sizeof(long long) * (size_t)n_lines, cudaMemcpyDeviceToHost, s));
        }
        OBJ_CK(cudaStreamSynchronize(s));
        if (h_err & ERR_INDEX) {
            *err = "index out of bounds: a face refers to a vertex / normal / texture coordinate that was not read before it";
            rc = RL_E_INVALID;
            goto done;
        }
        if (h_err & ERR_PRECISION) {
            *err = "a number has more significant digits than the device parser converts exactly (19 digits, |exp10| <= 60)";
            rc = RL_E_UNSUPPORTED;
            goto done;
        }
        info->n_triangles = ntri;
        if (ng == 0) {
            *out = raw;
            raw = ObjMesh{};
        } else {
            struct Grp { std::string name; bool is_default; int src, count; };
            std::vector<Grp> order;  // insertion order; a re-inserted key keeps its slot and takes the new value
            auto insert = [&](const std::string& name, bool dflt, int src, int count) {
                for (Grp& g : order)
                    if (g.is_default == dflt && g.name == name) { g.src = src; g.count = count; return; }
                order.push_back(Grp{name, dflt, src, count});
            };
            std::string cur;
            bool cur_default = true;
            int seg_src = 0;
            for (int li = 0; li < n_lines; li++) {
                if (h_kind[(size_t)li] != LK_G) continue;
                insert(cur, cur_default, seg_src, h_tri[(size_t)li] - seg_src);
                seg_src = h_tri[(size_t)li];
                long long b = h_start[(size_t)li], e = li + 1 < n_lines ? h_start[(size_t)li + 1] : n;
                std::string line(text + b, text + e);
                while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
                std::string tail = line.substr(line.find(' ') + 1);
                size_t tb = 0, te = tail.size();
                auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
                while (tb < te && ws(tail[tb])) tb++;
                while (te > tb && ws(tail[te - 1])) te--;
                cur = tail.substr(tb, te - tb);
                cur_default = false;
            }
            insert(cur, cur_default, seg_src, ntri - seg_src);
            int dst = 0;
            for (const Grp& g : order) {
                if (g.count > 0) segs.push_back(Segment{g.src, g.count, dst});
                dst += g.count;
            }
            out->n_triangles = dst;
            info->n_triangles = dst;
            OBJ_CK(dev_alloc(&out->tri_p, 9 * (size_t)dst));
            OBJ_CK(dev_alloc(&out->tri_n, 9 * (size_t)dst));
            OBJ_CK(dev_alloc(&out->tri_uv, 6 * (size_t)dst));
            OBJ_CK(dev_alloc(&out->tri_flags, (size_t)dst));
            if (!segs.empty()) {
                OBJ_CK(dev_alloc(&d_segs, segs.size()));
                OBJ_CK(cudaMemcpyAsync(d_segs, segs.data(), sizeof(Segment) * segs.size(), cudaMemcpyHostToDevice, s));
                dim3 grid(64, (unsigned)segs.size());
                k_gather_segments<<<grid, 256, 0, s>>>(d_segs, (int)segs.size(), raw, *out);
                launches++;
                OBJ_CK(cudaStreamSynchronize(s));
            }
        }
        if (out->n_triangles > 0) {
            double* d_b = nullptr;
            OBJ_CK(dev_alloc(&d_b, 6));
            k_mesh_bounds<<<1, 256, 0, s>>>(out->tri_p, 3 * out->n_triangles, d_b);
            launches++;
            cudaError_t ce = cudaMemcpyAsync(info->bounds, d_b, 6 * sizeof(double), cudaMemcpyDeviceToHost, s);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
            cudaFree(d_b);
            OBJ_CK(ce);
        }
        info->kernel_launches = launches;
    }
done:
    if (rc != RL_OK) obj_mesh_free(out);
    obj_mesh_free(&raw);
    void* tmp[] = {d_text, flag, pos, scratch, derr, start, kind, is_v, is_vn, is_vt, is_g, n_tris, vals, verts, norms, tex, d_segs};
    for (void* p : tmp)
        if (p) cudaFree(p);
    return rc;
}

}  // namespace rl
