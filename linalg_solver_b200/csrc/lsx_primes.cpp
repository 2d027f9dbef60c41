// Host-side number theory for liblsx: the 31-bit prime table, Montgomery constants and the
// Hadamard-bound -> (primes, limbs) planning used by every lsx_plan_* entry point.
#include <cmath>

#include "lsx_internal.h"

static uint32_t mulmod_u32(uint32_t a, uint32_t b, uint32_t n) { return (uint32_t)((uint64_t)a * b % n); }

static uint32_t powmod_u32(uint32_t a, uint32_t e, uint32_t n) {
    uint32_t r = 1 % n;
    a %= n;
    while (e) {
        if (e & 1) r = mulmod_u32(r, a, n);
        a = mulmod_u32(a, a, n);
        e >>= 1;
    }
    return r;
}

// Deterministic Miller-Rabin: bases {2,3,5,7} are exact below 3,215,031,751 > 2^31.
bool lsx_is_prime_u32(uint32_t n) {
    if (n < 2) return false;
    static const uint32_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint32_t q : small) {
        if (n % q == 0) return n == q;
    }
    uint32_t d = n - 1;
    int s = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++s;
    }
    for (uint32_t a : {2u, 3u, 5u, 7u}) {
        uint32_t x = powmod_u32(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (int i = 1; i < s; ++i) {
            x = mulmod_u32(x, x, n);
            if (x == n - 1) {
                composite = false;
                break;
            }
        }
        if (composite) return false;
    }
    return true;
}

// Primes below 2^31 in descending order, starting with 2^31 - 1.
void lsx_fill_prime_table(std::vector<uint32_t>& out, int count) {
    out.clear();
    uint32_t n = 0x7fffffffu;
    while ((int)out.size() < count) {
        if (lsx_is_prime_u32(n)) out.push_back(n);
        n -= 2;
    }
}

uint32_t lsx_inv_mod(uint32_t a, uint32_t p) {
    // extended Euclid on signed 64-bit
    int64_t t = 0, newt = 1, r = p, newr = a % p;
    while (newr != 0) {
        int64_t q = r / newr;
        int64_t tmp = t - q * newt;
        t = newt;
        newt = tmp;
        tmp = r - q * newr;
        r = newr;
        newr = tmp;
    }
    if (t < 0) t += p;
    return (uint32_t)t;
}

PrimeRec lsx_make_prime_rec(uint32_t p) {
    PrimeRec r;
    r.p = p;
    // Newton iteration for p^{-1} mod 2^32 (p odd)
    uint32_t inv = p;
    for (int i = 0; i < 5; ++i) inv *= 2u - p * inv;
    r.pinv = 0u - inv;
    uint64_t R = (uint64_t)1 << 32;
    r.one = (uint32_t)(R % p);
    r.r2 = (uint32_t)((uint64_t)r.one * r.one % p);
    return r;
}

// Every output integer of an elimination of [A|B] (pivots only in the first `bar` columns) is a
// minor of [A|B]: the common denominator is an r x r minor of A (r = rank <= min(m, bar,
// max_rank)); numerators in pivot rows are r x r minors with at most one column of B; entries
// of non-pivot rows are (r+1) x (r+1) bordered minors with exactly one column of B.  Hadamard:
// |minor| <= prod of column 2-norms.  A unit column (right block = identity) lowers the size by
// one through cofactor expansion.
double lsx_log2_minor_bound(int m, int bar, bool has_right, int64_t a_abs, int64_t b_abs,
                            bool right_identity, int max_rank) {
    double a = (double)(a_abs < 1 ? 1 : a_abs);
    double b = (double)(b_abs < 1 ? 1 : b_abs);
    int r = m < bar ? m : bar;
    if (max_rank > 0 && max_rank < r) r = max_rank;
    double best = r > 0 ? r * (0.5 * std::log2((double)r) + std::log2(a)) : 0.0;
    if (has_right) {
        int s = (r + 1 < m) ? r + 1 : m;
        double cand;
        if (right_identity) {
            int t = s - 1;
            cand = t > 0 ? t * (0.5 * std::log2((double)t) + std::log2(a)) : 0.0;
        } else {
            cand = (s - 1) * (0.5 * std::log2((double)s) + std::log2(a)) + 0.5 * std::log2((double)s) +
                   std::log2(b);
        }
        if (cand > best) best = cand;
    }
    return best;
}

// Table primes are all > 2^30.999 (the first 2048 primes below 2^31 are within 2^16 of it).
void lsx_bits_to_plan(double log2_bound, int* n_primes, int* limbs) {
    const double need = log2_bound + 1.0 + 1e-6;   // sign bit + floating-point slack
    int K = (int)std::ceil(need / 30.999);
    int L = (int)std::ceil(need / 32.0);
    *n_primes = K < 1 ? 1 : K;
    *limbs = L < 1 ? 1 : L;
}
