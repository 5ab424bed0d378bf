// mt_jump.h -- GF(2) jump-ahead polynomials for MT19937 (host side; see mt_jump.cpp)
#pragma once
#include <stdint.h>

namespace gsage {
// t^steps mod phi as 624 x 32-bit words (bit i = coefficient of t^i); 0 on success
int mt_jump_poly(uint64_t steps, uint32_t* poly_out);
// polys_out[i] = t^(stride*(i+1)) mod phi, i in [0, count)
int mt_jump_poly_series(uint64_t stride, int count, uint32_t* polys_out);
// out = g(F) window   (CPU reference of the GPU jump kernel)
void mt_jump_apply_host(const uint32_t* window, const uint32_t* poly, uint32_t* out);
}
