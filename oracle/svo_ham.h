/* svo_ham.h — 256-bit Hamming distance shared by the oracle's translation units (test infrastructure). */
#ifndef SVO_HAM_H
#define SVO_HAM_H
#include <stdint.h>
#include <string.h>
/* Same distance with the popcnt instruction: what -O3 -march=native makes of the loops that
 * call DescriptorDistance; used by the bulk matchers below so the CPU baseline is not
 * handicapped by the SWAR form (tests check both agree). */
static inline int ham256(const uint8_t *a, const uint8_t *b)
{
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}
#endif
