// mt_jump.cpp -- host-side GF(2) polynomial arithmetic for MT19937 jump-ahead.
//
// The reference consumes ONE sequential MT19937 stream (numpy's global RandomState,
// /root/reference/nn_modules.py:88).  To refill that stream on a GPU at tens of Gwords/s the generator is
// split into lanes that each start `J` words further down the same stream.  Jumping a state ahead by J
// words is the classic polynomial method (Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008):
//
//     g_J(t) = t^J mod phi(t),      F^J s = g_J(F) s = XOR_{i : g_J[i] = 1} F^i s,
//
// with phi the characteristic polynomial (degree 19937) of the one-word transition F.  Because the state
// is a 624-word window of the stream, F^i s is simply the window starting i words later, so the jumped
// window is an XOR of shifted windows of the next 19937+624 words -- embarrassingly parallel on the GPU
// (mt19937.cu: mt_jump_kernel).  This file computes phi (Berlekamp-Massey on 2*19937 output bits) and the
// jump polynomials; tests/test_mt_jump.py checks them on the CPU against numpy's own stream.
#include "mt_jump.h"

#include <string.h>
#include <algorithm>
#include <vector>

namespace gsage {

static const int DEG = 19937;
static const int W64 = (2 * DEG + 64) / 64 + 1;     // enough 64-bit words for degree < 2*DEG products

typedef std::vector<uint64_t> Poly;                  // bit i of word i/64 = coefficient of t^i

static inline int get(const Poly& p, int i) { return (int)((p[i >> 6] >> (i & 63)) & 1u); }
static inline void flip(Poly& p, int i) { p[i >> 6] ^= (uint64_t)1 << (i & 63); }

static inline uint32_t mix(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

// the infinite word sequence continuing a 624-word window: x[n+624] = x[n+397] ^ mix(x[n], x[n+1])
static void extend(const uint32_t* window, size_t total, std::vector<uint32_t>& x) {
    x.resize(total);
    memcpy(x.data(), window, sizeof(uint32_t) * 624);
    for (size_t n = 0; n + 624 < total; ++n) x[n + 624] = x[n + 397] ^ mix(x[n], x[n + 1]);
}

// dst ^= src << shift   (polynomials over GF(2), word-packed)
static void xor_shifted(Poly& dst, const Poly& src, int src_words, int shift) {
    const int ws = shift >> 6, bs = shift & 63;
    if (bs == 0) {
        for (int i = 0; i < src_words; ++i) dst[i + ws] ^= src[i];
    } else {
        for (int i = 0; i < src_words; ++i) {
            dst[i + ws] ^= src[i] << bs;
            dst[i + ws + 1] ^= src[i] >> (64 - bs);
        }
    }
}

// Berlekamp-Massey over GF(2): connection polynomial C (C[0] = 1) of the shortest LFSR generating s[0..n):
// s[i] = XOR_{j=1..L} C[j] s[i-j].  Word-packed: R holds the reversed history (bit j = s[i-j]) so the
// discrepancy is parity(C & R).
static int berlekamp_massey(const std::vector<uint8_t>& s, Poly& C) {
    const int n = (int)s.size();
    const int words = n / 64 + 3;
    Poly B(words, 0), T, R(words, 0);
    C.assign(words, 0);
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    for (int i = 0; i < n; ++i) {
        uint64_t carry = s[i];                               // R <<= 1; R[0] = s[i]
        const int used = std::min(words, i / 64 + 2);
        for (int w = 0; w < used; ++w) {
            const uint64_t next = R[w] >> 63;
            R[w] = (R[w] << 1) | carry;
            carry = next;
        }
        uint64_t acc = 0;
        const int cw = L / 64 + 1;
        for (int w = 0; w < cw; ++w) acc ^= C[w] & R[w];
        if (!__builtin_parityll(acc)) { ++m; continue; }
        const int src_words = std::min(words - (m >> 6) - 1, i / 64 + 2);
        if (2 * L <= i) {
            T = C;
            xor_shifted(C, B, src_words, m);
            L = i + 1 - L;
            B = T;
            m = 1;
        } else {
            xor_shifted(C, B, src_words, m);
            ++m;
        }
    }
    return L;
}

struct Phi {
    bool ready = false;
    Poly phi;          // degree DEG, monic
};
static Phi g_phi;

static bool compute_phi() {
    if (g_phi.ready) return true;
    // any non-degenerate seed; bit 0 of the untempered words is a linear functional of the state
    uint32_t key[624];
    key[0] = 19650218u;
    for (int i = 1; i < 624; ++i) key[i] = 1812433253u * (key[i - 1] ^ (key[i - 1] >> 30)) + (uint32_t)i;
    std::vector<uint32_t> x;
    extend(key, 624 + 2 * DEG + 8, x);
    std::vector<uint8_t> s(2 * DEG + 4);
    for (size_t i = 0; i < s.size(); ++i) s[i] = (uint8_t)(x[624 + i] & 1u);
    Poly C;
    const int L = berlekamp_massey(s, C);
    if (L != DEG) return false;
    // phi(t) = t^L * C(1/t):  phi_k = C_{L-k}
    g_phi.phi.assign(W64, 0);
    for (int k = 0; k <= DEG; ++k)
        if (get(C, DEG - k)) flip(g_phi.phi, k);
    g_phi.ready = get(g_phi.phi, DEG) == 1;
    return g_phi.ready;
}

// p (degree < 2*DEG) mod phi, in place; result degree < DEG
static void reduce(Poly& p) {
    const Poly& phi = g_phi.phi;
    const int phi_words = DEG / 64 + 1;
    for (int i = 2 * DEG; i >= DEG; --i)
        if (get(p, i)) xor_shifted(p, phi, phi_words, i - DEG);
}

static void mul_mod(const Poly& a, const Poly& b, Poly& out) {
    Poly acc(W64 + 2, 0);
    const int words = DEG / 64 + 1;
    for (int i = 0; i < DEG; ++i)
        if (get(b, i)) xor_shifted(acc, a, words, i);
    reduce(acc);
    out.assign(W64 + 2, 0);
    for (int w = 0; w < words; ++w) out[w] = acc[w];
    // clear anything at or above DEG (reduce leaves none, but keep the invariant explicit)
    for (int i = DEG; i < (words * 64); ++i)
        if (get(out, i)) flip(out, i);
}

static void pow_t_mod(uint64_t e, Poly& out) {
    // t^e mod phi by left-to-right square and multiply-by-t
    Poly r(W64 + 2, 0);
    r[0] = 1;                                      // t^0
    int top = 63;
    while (top >= 0 && !((e >> top) & 1)) --top;
    for (int b = top; b >= 0; --b) {
        Poly sq;
        mul_mod(r, r, sq);
        r = sq;
        if ((e >> b) & 1) {                        // multiply by t: shift left one, reduce
            Poly sh(W64 + 2, 0);
            xor_shifted(sh, r, DEG / 64 + 1, 1);
            reduce(sh);
            r = sh;
        }
    }
    out = r;
}

int mt_jump_poly(uint64_t steps, uint32_t* poly_out /* 624 words */) {
    if (!compute_phi()) return -1;
    Poly g;
    pow_t_mod(steps, g);
    for (int w = 0; w < 312; ++w) {
        poly_out[2 * w] = (uint32_t)(g[w] & 0xFFFFFFFFu);
        poly_out[2 * w + 1] = (uint32_t)(g[w] >> 32);
    }
    return 0;
}

int mt_jump_poly_series(uint64_t stride, int count, uint32_t* polys_out /* count x 624 words: t^(stride*(i+1)) */) {
    if (!compute_phi()) return -1;
    Poly g1, cur;
    pow_t_mod(stride, g1);
    cur = g1;
    for (int i = 0; i < count; ++i) {
        if (i > 0) {
            Poly next;
            mul_mod(cur, g1, next);
            cur = next;
        }
        uint32_t* dst = polys_out + (size_t)i * 624;
        for (int w = 0; w < 312; ++w) {
            dst[2 * w] = (uint32_t)(cur[w] & 0xFFFFFFFFu);
            dst[2 * w + 1] = (uint32_t)(cur[w] >> 32);
        }
    }
    return 0;
}

void mt_jump_apply_host(const uint32_t* window, const uint32_t* poly, uint32_t* out) {
    std::vector<uint32_t> x;
    extend(window, 624 + DEG + 1, x);
    memset(out, 0, sizeof(uint32_t) * 624);
    for (int i = 0; i < DEG; ++i)
        if ((poly[i >> 5] >> (i & 31)) & 1u)
            for (int j = 0; j < 624; ++j) out[j] ^= x[i + j];
}

}  // namespace gsage

extern "C" {

// test hooks (CPU only): tests/test_mt_jump.py pins the polynomial arithmetic against numpy's stream
int gsage_mt_jump_poly(uint64_t steps, uint32_t* poly_out) { return gsage::mt_jump_poly(steps, poly_out); }
void gsage_mt_jump_apply_host(const uint32_t* window, const uint32_t* poly, uint32_t* out) {
    gsage::mt_jump_apply_host(window, poly, out);
}

}  // extern "C"
