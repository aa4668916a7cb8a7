// oracle/terrain.h — TEST INFRASTRUCTURE (CPU oracle side).  Not part of the product.
//
// Integer-only value-noise height field used as the synthetic "noise terrain" scene
// (the reference ships no terrain generator, BASELINE.json configs 2-4 need one).
// The definition is normative (DESIGN.md §"Terrain editor"); the CUDA product re-implements it
// independently in vkhashdag_b200/csrc/editors.cuh and is parity-checked against this file.
//
//   lattice(o,ix,iz) = top 16 bits of fmix32( fmix32( fmix32(seed + o*0x9E3779B9) ^ ix*0x85EBCA6B ) ^ iz*0xC2B2AE35 )
//   octave o: cell = 2^c, c = first_cell_bits - 2o, amplitude a = first_amplitude >> 2o
//             r(x,z) = bilinear integer interpolation of the 4 lattice values, in [0,65535]
//   height(x,z) = base + sum_o (r_o(x,z) * a_o) >> 16 ;  voxel solid iff y < height(x,z)
//   optional patch: with extent_bits = e != 0 only columns x,z < 2^e carry terrain (rest of the world stays empty)
#pragma once
#include <cstdint>

namespace terrain {

struct Params {
	uint32_t seed, base, first_cell_bits, octaves, first_amplitude;
	uint32_t extent_bits = 0; // != 0: the terrain only exists for x,z < 2^extent_bits (a patch of a larger world)
};

inline Params from_desc(uint32_t aux, const uint32_t p0[3], const uint32_t p1[3]) {
	return Params{aux, p0[0], p0[1], p0[2], p1[0], p1[1]};
}
// footprint [lx,lx+2^k) x [lz,lz+2^k): 0 = outside the patch, 1 = partly inside, 2 = fully inside
inline int extent_class(const Params &p, uint32_t lx, uint32_t lz, uint32_t k) {
	if (p.extent_bits == 0)
		return 2;
	const uint64_t e = 1ull << p.extent_bits, s = 1ull << k;
	if (lx >= e || lz >= e)
		return 0;
	return (lx + s <= e && lz + s <= e) ? 2 : 1;
}
inline bool in_extent(const Params &p, uint32_t x, uint32_t z) {
	return p.extent_bits == 0 || ((x >> p.extent_bits) == 0 && (z >> p.extent_bits) == 0);
}

inline uint32_t fmix32(uint32_t h) {
	h ^= h >> 16;
	h *= 0x85ebca6bu;
	h ^= h >> 13;
	h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}

inline uint32_t lattice(const Params &p, uint32_t o, uint32_t ix, uint32_t iz) {
	uint32_t h = fmix32(p.seed + o * 0x9E3779B9u);
	h = fmix32(h ^ (ix * 0x85EBCA6Bu));
	h = fmix32(h ^ (iz * 0xC2B2AE35u));
	return h >> 16;
}

// bilinear value inside the cell (ix,iz) at offsets fx,fz in [0, 2^c] (inclusive upper end allowed)
inline uint32_t bilerp(uint32_t v00, uint32_t v10, uint32_t v01, uint32_t v11, uint32_t fx, uint32_t fz, uint32_t c) {
	uint64_t S = 1ull << c;
	uint64_t a = uint64_t(v00) * (S - fx) + uint64_t(v10) * fx;
	uint64_t b = uint64_t(v01) * (S - fx) + uint64_t(v11) * fx;
	return uint32_t((a * (S - fz) + b * fz) >> (2 * c));
}

inline bool octave_active(const Params &p, uint32_t o) { return o < p.octaves && p.first_cell_bits >= 2 * o; }

inline uint32_t height(const Params &p, uint32_t x, uint32_t z) {
	uint32_t h = p.base;
	for (uint32_t o = 0; octave_active(p, o); ++o) {
		uint32_t c = p.first_cell_bits - 2 * o, amp = p.first_amplitude >> (2 * o);
		uint32_t ix = x >> c, iz = z >> c, m = (1u << c) - 1u;
		uint32_t r = bilerp(lattice(p, o, ix, iz), lattice(p, o, ix + 1, iz), lattice(p, o, ix, iz + 1),
		                    lattice(p, o, ix + 1, iz + 1), x & m, z & m, c);
		h += uint32_t((uint64_t(r) * amp) >> 16);
	}
	return h;
}

// Conservative [hmin, hmax] of height() over the footprint [lx,lx+2^k) x [lz,lz+2^k), lx,lz multiples of 2^k.
inline void height_bounds(const Params &p, uint32_t lx, uint32_t lz, uint32_t k, uint32_t &hmin, uint32_t &hmax) {
	hmin = hmax = p.base;
	for (uint32_t o = 0; octave_active(p, o); ++o) {
		uint32_t c = p.first_cell_bits - 2 * o, amp = p.first_amplitude >> (2 * o);
		uint32_t rmin, rmax;
		if (k <= c) { // footprint inside one lattice cell: extremes of a bilinear patch are at its corners
			uint32_t ix = lx >> c, iz = lz >> c, m = (1u << c) - 1u;
			uint32_t v00 = lattice(p, o, ix, iz), v10 = lattice(p, o, ix + 1, iz), v01 = lattice(p, o, ix, iz + 1),
			         v11 = lattice(p, o, ix + 1, iz + 1);
			uint32_t fx0 = lx & m, fz0 = lz & m, fx1 = fx0 + (1u << k), fz1 = fz0 + (1u << k);
			uint32_t r0 = bilerp(v00, v10, v01, v11, fx0, fz0, c), r1 = bilerp(v00, v10, v01, v11, fx1, fz0, c),
			         r2 = bilerp(v00, v10, v01, v11, fx0, fz1, c), r3 = bilerp(v00, v10, v01, v11, fx1, fz1, c);
			rmin = r0 < r1 ? r0 : r1;
			rmax = r0 > r1 ? r0 : r1;
			uint32_t lo = r2 < r3 ? r2 : r3, hi = r2 > r3 ? r2 : r3;
			rmin = lo < rmin ? lo : rmin;
			rmax = hi > rmax ? hi : rmax;
		} else if (k - c <= 2) { // up to 4x4 cells: bounded by the covered lattice values
			uint32_t n = 1u << (k - c), ix = lx >> c, iz = lz >> c;
			rmin = 0xFFFFu, rmax = 0;
			for (uint32_t j = 0; j <= n; ++j)
				for (uint32_t i = 0; i <= n; ++i) {
					uint32_t v = lattice(p, o, ix + i, iz + j);
					rmin = v < rmin ? v : rmin;
					rmax = v > rmax ? v : rmax;
				}
		} else {
			rmin = 0, rmax = 0xFFFFu;
		}
		hmin += uint32_t((uint64_t(rmin) * amp) >> 16);
		hmax += uint32_t((uint64_t(rmax) * amp) >> 16);
	}
}

} // namespace terrain
