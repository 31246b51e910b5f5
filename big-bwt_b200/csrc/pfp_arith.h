/* pfp_arith.h -- exact integer arithmetic of the trigger scan, shared by device and host.
 *
 * The reference rolls a Karp-Rabin fingerprint of the last w bytes, base 256, modulo
 * PW = 1999999973 (newscan.cpp:172,194-202), with two 64-bit `%` per byte, and cuts a
 * phrase when `hash % p == 0` (newscan.cpp:344,367).  A 64-bit `%` costs ~13 integer
 * instructions on the GPU, so the same VALUES are computed here division-free:
 *
 *   roll  : t = 256*h + c_in + c_out*NEGW            (t < 2^41, NEGW = -256^w mod PW)
 *   reduce: q = mulhi(t >> 20, floor(2^52/PW))       (q in {floor(t/PW)-1, floor(t/PW)})
 *           r = lo32(t) - q*PW ; r = min(r, r-PW)    (unsigned wrap makes the min exact)
 *   test  : p = 2^k * q_odd ;  r % p == 0  <=>  ror32(r * inv(q_odd), k) <= (2^32-1)/p
 *
 * tests/test_arith.py checks every formula against plain `%` on the host.
 */
#ifndef PFP_ARITH_H
#define PFP_ARITH_H
#include <stdint.h>

#if defined(__CUDACC__)
#define PFP_HD __host__ __device__ __forceinline__
#else
#define PFP_HD static inline
#endif

#define PFP_PW 1999999973u          /* newscan.cpp:172 */
#define PFP_PW_MAGIC 2251799u       /* floor(2^52 / PFP_PW) */
#define PFP_DOLLAR 2u               /* utils.h:5 */
#define PFP_END_OF_WORD 1u          /* utils.h:6 */
#define PFP_END_OF_DICT 0u          /* utils.h:7 */
#define PFP_IBYTES 5                /* utils.h:10 */

typedef struct pfp_scan_consts {
    uint32_t w, p;
    uint32_t negw;     /* (PW - 256^w mod PW) mod PW : contribution of the byte leaving      */
    uint32_t pinv;     /* inverse of the odd part of p modulo 2^32                           */
    uint32_t pshift;   /* number of trailing zero bits of p                                  */
    uint32_t plimit;   /* floor((2^32-1) / p)                                                */
} pfp_scan_consts;

/* t mod PW for t < 2^41 */
PFP_HD uint32_t pfp_reduce_pw(uint64_t t) {
    uint32_t x = (uint32_t)(t >> 20);
#if defined(__CUDA_ARCH__)
    uint32_t q = __umulhi(x, PFP_PW_MAGIC);
#else
    uint32_t q = (uint32_t)(((uint64_t)x * PFP_PW_MAGIC) >> 32);
#endif
    uint32_t r = (uint32_t)t - q * PFP_PW;
    uint32_t r2 = r - PFP_PW;
    return r2 < r ? r2 : r;
}

/* one Horner step without a byte leaving (window still filling) */
PFP_HD uint32_t pfp_push(uint32_t h, uint32_t c_in) {
    return pfp_reduce_pw(((uint64_t)h << 8) + c_in);
}

/* one rolling step: byte c_out leaves the w-window, c_in enters */
PFP_HD uint32_t pfp_roll(uint32_t h, uint32_t c_in, uint32_t c_out, uint32_t negw) {
    uint64_t t = ((uint64_t)h << 8) + (uint64_t)c_out * negw + c_in;
    return pfp_reduce_pw(t);
}

/* r % p == 0 for r < 2^32 */
PFP_HD int pfp_is_trigger(uint32_t r, uint32_t pinv, uint32_t pshift, uint32_t plimit) {
    uint32_t v = r * pinv;
#if defined(__CUDA_ARCH__)
    v = __funnelshift_r(v, v, pshift);
#else
    v = pshift ? ((v >> pshift) | (v << (32 - pshift))) : v;
#endif
    return v <= plimit;
}

static inline pfp_scan_consts pfp_make_scan_consts(uint32_t w, uint32_t p) {
    pfp_scan_consts c;
    c.w = w; c.p = p;
    uint64_t pw = 1;
    for (uint32_t i = 0; i < w; i++) pw = (pw * 256) % PFP_PW;
    c.negw = (uint32_t)((PFP_PW - pw) % PFP_PW);
    uint32_t k = 0, q = p;
    while ((q & 1u) == 0) { q >>= 1; k++; }
    uint32_t inv = q;                       /* Newton: inv *= 2 - q*inv doubles the correct bits */
    for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
    c.pinv = inv; c.pshift = k; c.plimit = 0xFFFFFFFFu / p;
    return c;
}

#endif
