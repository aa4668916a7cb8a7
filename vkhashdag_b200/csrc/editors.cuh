// editors.cuh — device predicates selected by hd_edit_desc::kind.
//
// Semantics follow the reference's editor structs (src/main.cpp:32-150): EditNode classifies an octree node by
// its integer voxel-space bounds (NodeCoord::GetLower/UpperBoundAtLevel, include/hashdag/NodeCoord.hpp:86-93),
// EditVoxel decides one voxel.  All arithmetic is exact integer (i64/u64), so CPU and GPU agree bit for bit.
// The terrain generator is the synthetic scene of SURVEY.md §8d cfg2 (normative definition in DESIGN.md).
#pragma once
#include "common.cuh"

namespace hd {

enum EditType : uint32_t { kNotAffected = 0, kProceed = 1, kFill = 2, kClear = 3 }; // include/hashdag/Editor.hpp:18

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
	h ^= h >> 16;
	h *= 0x85ebca6bu;
	h ^= h >> 13;
	h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}

struct TerrainOctave {
	uint32_t cell_bits, amp, hseed;
};

__device__ __forceinline__ uint32_t terrain_lattice(uint32_t hseed, uint32_t ix, uint32_t iz) {
	return fmix32(fmix32(hseed ^ (ix * 0x85EBCA6Bu)) ^ (iz * 0xC2B2AE35u)) >> 16;
}
__device__ __forceinline__ uint32_t terrain_bilerp(uint32_t v00, uint32_t v10, uint32_t v01, uint32_t v11, uint32_t fx,
                                                   uint32_t fz, uint32_t c) {
	const uint64_t S = 1ull << c;
	const uint64_t a = uint64_t(v00) * (S - fx) + uint64_t(v10) * fx;
	const uint64_t b = uint64_t(v01) * (S - fx) + uint64_t(v11) * fx;
	return uint32_t((a * (S - fz) + b * fz) >> (2 * c));
}
__device__ __forceinline__ bool terrain_octave(const hd_edit_desc &d, uint32_t o, TerrainOctave &out) {
	if (o >= d.p0[2] || d.p0[1] < 2 * o)
		return false;
	out.cell_bits = d.p0[1] - 2 * o;
	out.amp = d.p1[0] >> (2 * o);
	out.hseed = fmix32(d.aux + o * 0x9E3779B9u);
	return true;
}

__device__ inline uint32_t terrain_height(const hd_edit_desc &d, uint32_t x, uint32_t z) {
	uint32_t h = d.p0[0];
	TerrainOctave oc;
	for (uint32_t o = 0; terrain_octave(d, o, oc); ++o) {
		const uint32_t c = oc.cell_bits, ix = x >> c, iz = z >> c, m = (1u << c) - 1u;
		const uint32_t r = terrain_bilerp(terrain_lattice(oc.hseed, ix, iz), terrain_lattice(oc.hseed, ix + 1, iz),
		                                  terrain_lattice(oc.hseed, ix, iz + 1), terrain_lattice(oc.hseed, ix + 1, iz + 1),
		                                  x & m, z & m, c);
		h += uint32_t((uint64_t(r) * oc.amp) >> 16);
	}
	return h;
}

// conservative height range over the footprint [lx,lx+2^k) x [lz,lz+2^k)
__device__ inline void terrain_bounds(const hd_edit_desc &d, uint32_t lx, uint32_t lz, uint32_t k, uint32_t &hmin,
                                      uint32_t &hmax) {
	hmin = hmax = d.p0[0];
	TerrainOctave oc;
	for (uint32_t o = 0; terrain_octave(d, o, oc); ++o) {
		const uint32_t c = oc.cell_bits;
		uint32_t rmin, rmax;
		if (k <= c) {
			const uint32_t ix = lx >> c, iz = lz >> c, m = (1u << c) - 1u;
			const uint32_t v00 = terrain_lattice(oc.hseed, ix, iz), v10 = terrain_lattice(oc.hseed, ix + 1, iz),
			               v01 = terrain_lattice(oc.hseed, ix, iz + 1), v11 = terrain_lattice(oc.hseed, ix + 1, iz + 1);
			const uint32_t fx0 = lx & m, fz0 = lz & m, fx1 = fx0 + (1u << k), fz1 = fz0 + (1u << k);
			const uint32_t r0 = terrain_bilerp(v00, v10, v01, v11, fx0, fz0, c),
			               r1 = terrain_bilerp(v00, v10, v01, v11, fx1, fz0, c),
			               r2 = terrain_bilerp(v00, v10, v01, v11, fx0, fz1, c),
			               r3 = terrain_bilerp(v00, v10, v01, v11, fx1, fz1, c);
			rmin = min(min(r0, r1), min(r2, r3));
			rmax = max(max(r0, r1), max(r2, r3));
		} else if (k - c <= 2) {
			const uint32_t n = 1u << (k - c), ix = lx >> c, iz = lz >> c;
			rmin = 0xFFFFu, rmax = 0u;
			for (uint32_t j = 0; j <= n; ++j)
				for (uint32_t i = 0; i <= n; ++i) {
					const uint32_t v = terrain_lattice(oc.hseed, ix + i, iz + j);
					rmin = min(rmin, v), rmax = max(rmax, v);
				}
		} else {
			rmin = 0u, rmax = 0xFFFFu;
		}
		hmin += uint32_t((uint64_t(rmin) * oc.amp) >> 16);
		hmax += uint32_t((uint64_t(rmax) * oc.amp) >> 16);
	}
}

// EditNode of the EIGHT children of the node at (x, y, z) in one go: children sit at (2x + cx, 2y + cy, 2z + cz) on a level
// whose nodes span 2^bits voxels.  Returns two bits per child (child c = cx | cy << 1 | cz << 2 at bits 2c, 2c + 1) holding the
// EditType edit_node() returns for it.  The children share three planes per axis, so the per-axis terms (squared distances,
// inside / outside flags) are formed once for the two halves instead of once per child: ~8 instructions per child instead
// of ~45.  Same integer arithmetic, value for value: the 32-bit sphere path is taken when the PARENT's box allows it (a
// superset of every child's box), the 64-bit one otherwise — both are exact, so the choice never changes a result.
// Terrain edits are not handled here (callers use edit_node for batches that hold one).
__device__ inline uint32_t edit_node8(const hd_edit_desc &d, uint32_t bits, uint32_t x, uint32_t y, uint32_t z) {
	const uint32_t s = 1u << bits;
	const uint32_t L[3] = {(x << 1) << bits, (y << 1) << bits, (z << 1) << bits};
	uint32_t out = 0u;
	switch (d.kind) {
	case HD_EDIT_AABB_FILL: { // main.cpp:35-46
		bool outside[3][2], inside[3][2];
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const uint32_t M = L[a] + s, H = M + s;
			outside[a][0] = M <= d.p0[a] || L[a] >= d.p1[a], inside[a][0] = L[a] >= d.p0[a] && M <= d.p1[a];
			outside[a][1] = H <= d.p0[a] || M >= d.p1[a], inside[a][1] = M >= d.p0[a] && H <= d.p1[a];
		}
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			const int hx = c & 1, hy = (c >> 1) & 1, hz = c >> 2;
			const bool o = outside[0][hx] || outside[1][hy] || outside[2][hz];
			const bool in = inside[0][hx] && inside[1][hy] && inside[2][hz];
			out |= uint32_t(o ? kNotAffected : (in ? kFill : kProceed)) << (2 * c);
		}
		return out;
	}
	case HD_EDIT_SPHERE_FILL:
	case HD_EDIT_SPHERE_DIG: { // main.cpp:77-106
		const uint32_t term = d.kind == HD_EDIT_SPHERE_DIG ? kClear : kFill;
		int32_t dl[3], dm[3], dh[3];
		uint32_t far = 0u;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			dl[a] = int32_t(L[a] - d.p0[a]), dm[a] = int32_t(L[a] + s - d.p0[a]), dh[a] = int32_t(L[a] + 2u * s - d.p0[a]);
			far = max(far, max(uint32_t(abs(dl[a])), uint32_t(abs(dh[a])))); // |dm| lies between them
		}
		if (far < 37837u && bits < 22u) { // three squares below 2^32 / 3 each: the sums cannot wrap
			uint32_t mx[3][2], mn[3][2];
#pragma unroll
			for (int a = 0; a < 3; ++a) {
				const uint32_t sl = uint32_t(dl[a] * dl[a]), sm = uint32_t(dm[a] * dm[a]), sh = uint32_t(dh[a] * dh[a]);
				mx[a][0] = max(sl, sm), mx[a][1] = max(sm, sh);
				mn[a][0] = dl[a] > 0 ? sl : (dm[a] < 0 ? sm : 0u);
				mn[a][1] = dm[a] > 0 ? sm : (dh[a] < 0 ? sh : 0u);
			}
#pragma unroll
			for (int c = 0; c < 8; ++c) {
				const int hx = c & 1, hy = (c >> 1) & 1, hz = c >> 2;
				const uint32_t big = mx[0][hx] + mx[1][hy] + mx[2][hz], small = mn[0][hx] + mn[1][hy] + mn[2][hz];
				out |= (uint64_t(big) <= d.r2 ? term : (uint64_t(small) > d.r2 ? uint32_t(kNotAffected) : uint32_t(kProceed))) << (2 * c);
			}
			return out;
		}
		uint64_t mx[3][2], mn[3][2];
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const long long l = (long long)L[a] - (long long)d.p0[a], m = l + (long long)s, h = m + (long long)s;
			const uint64_t sl = uint64_t(l * l), sm = uint64_t(m * m), sh = uint64_t(h * h);
			mx[a][0] = sl > sm ? sl : sm, mx[a][1] = sm > sh ? sm : sh;
			mn[a][0] = l > 0 ? sl : (m < 0 ? sm : 0ull);
			mn[a][1] = m > 0 ? sm : (h < 0 ? sh : 0ull);
		}
#pragma unroll
		for (int c = 0; c < 8; ++c) {
			const int hx = c & 1, hy = (c >> 1) & 1, hz = c >> 2;
			const uint64_t big = mx[0][hx] + mx[1][hy] + mx[2][hz], small = mn[0][hx] + mn[1][hy] + mn[2][hz];
			out |= (big <= d.r2 ? term : (small > d.r2 ? uint32_t(kNotAffected) : uint32_t(kProceed))) << (2 * c);
		}
		return out;
	}
	default:
		return 0u; // kNotAffected x 8
	}
}

// EditNode: node at `level` with integer position (x,y,z); bits = voxel_level - level.
template <bool kTerrain = true>
__device__ inline EditType edit_node(const hd_edit_desc &d, uint32_t bits, uint32_t x, uint32_t y, uint32_t z) {
	const uint32_t lb[3] = {x << bits, y << bits, z << bits};
	const uint32_t ub[3] = {(x + 1u) << bits, (y + 1u) << bits, (z + 1u) << bits};
	switch (d.kind) {
	case HD_EDIT_AABB_FILL: { // main.cpp:35-46
		bool outside = false, inside = true;
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			outside |= ub[i] <= d.p0[i] || lb[i] >= d.p1[i];
			inside &= lb[i] >= d.p0[i] && ub[i] <= d.p1[i];
		}
		return outside ? kNotAffected : (inside ? kFill : kProceed);
	}
	case HD_EDIT_SPHERE_FILL:
	case HD_EDIT_SPHERE_DIG: { // main.cpp:77-106
		// near the sphere (every |lo|, |hi| < 37 837) the sums stay below 2^32 and 32-bit arithmetic gives the same values
		const int32_t l0 = int32_t(lb[0] - d.p0[0]), l1 = int32_t(lb[1] - d.p0[1]), l2 = int32_t(lb[2] - d.p0[2]);
		const int32_t h0 = int32_t(ub[0] - d.p0[0]), h1 = int32_t(ub[1] - d.p0[1]), h2 = int32_t(ub[2] - d.p0[2]);
		const uint32_t far = max(max(max(uint32_t(abs(l0)), uint32_t(abs(h0))), max(uint32_t(abs(l1)), uint32_t(abs(h1)))),
		                         max(uint32_t(abs(l2)), uint32_t(abs(h2))));
		if (far < 37837u && bits < 22u) {
			const int32_t lo[3] = {l0, l1, l2}, hi[3] = {h0, h1, h2};
			uint32_t mx = 0, mn = 0;
#pragma unroll
			for (int i = 0; i < 3; ++i) {
				const uint32_t lo2 = uint32_t(lo[i] * lo[i]), hi2 = uint32_t(hi[i] * hi[i]);
				mx += lo2 > hi2 ? lo2 : hi2;
				if (lo[i] > 0)
					mn += lo2;
				if (hi[i] < 0)
					mn += hi2;
			}
			if (uint64_t(mx) <= d.r2)
				return d.kind == HD_EDIT_SPHERE_DIG ? kClear : kFill;
			return uint64_t(mn) > d.r2 ? kNotAffected : kProceed;
		}
		uint64_t max_n2 = 0, min_n2 = 0;
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			const long long lo = (long long)lb[i] - (long long)d.p0[i], hi = (long long)ub[i] - (long long)d.p0[i];
			const uint64_t lo2 = uint64_t(lo * lo), hi2 = uint64_t(hi * hi);
			max_n2 += lo2 > hi2 ? lo2 : hi2;
			if (lo > 0)
				min_n2 += lo2;
			if (hi < 0)
				min_n2 += hi2;
		}
		if (max_n2 <= d.r2)
			return d.kind == HD_EDIT_SPHERE_DIG ? kClear : kFill;
		return min_n2 > d.r2 ? kNotAffected : kProceed;
	}
	case HD_EDIT_TERRAIN_FILL: {
		if (!kTerrain)
			return kNotAffected;
		bool whole = true; // footprint fully inside the terrain patch (p1[1] = extent bits, 0 = whole world)
		if (d.p1[1] != 0u) {
			const uint64_t e = 1ull << d.p1[1], s = 1ull << bits;
			if (lb[0] >= e || lb[2] >= e)
				return kNotAffected;
			whole = lb[0] + s <= e && lb[2] + s <= e;
		}
		uint32_t hmin, hmax;
		terrain_bounds(d, lb[0], lb[2], bits, hmin, hmax);
		if (lb[1] >= hmax)
			return kNotAffected;
		if (whole && ub[1] <= hmin)
			return kFill;
		return kProceed;
	}
	}
	return kNotAffected;
}

// VoxelInRange (main.cpp:57-59,127-132).  kTerrain = false compiles the terrain generator out (kernels launched for
// batches without a terrain edit: fewer registers, more resident warps).
template <bool kTerrain = true>
__device__ inline bool voxel_in_range(const hd_edit_desc &d, uint32_t x, uint32_t y, uint32_t z) {
	switch (d.kind) {
	case HD_EDIT_AABB_FILL:
		return x >= d.p0[0] && y >= d.p0[1] && z >= d.p0[2] && x < d.p1[0] && y < d.p1[1] && z < d.p1[2];
	case HD_EDIT_SPHERE_FILL:
	case HD_EDIT_SPHERE_DIG: {
		const long long dx = (long long)x - (long long)d.p0[0], dy = (long long)y - (long long)d.p0[1],
		                dz = (long long)z - (long long)d.p0[2];
		return uint64_t(dx * dx + dy * dy + dz * dz) <= d.r2;
	}
	case HD_EDIT_TERRAIN_FILL:
		if (!kTerrain)
			return false;
		if (d.p1[1] != 0u && ((x >> d.p1[1]) != 0u || (z >> d.p1[1]) != 0u))
			return false;
		return y < terrain_height(d, x, z);
	}
	return false;
}

// Sphere VoxelInRange for the two voxels a lane owns in a leaf, (x, y, z) and (x, y, z + 2).  The reference compares
// dx^2 + dy^2 + dz^2 <= r2 in 64 bits (main.cpp:127-132); when every |d| < 37 837 the sum is below 2^32 and 32-bit
// arithmetic gives the same value (3 * 37 836^2 = 4 294 688 688 < 2^32) — always the case next to the sphere's surface
// for r < 37 000, i.e. at every leaf a brush reaches.  Larger distances take the 64-bit form.
__device__ __forceinline__ void sphere_in_range_pair(const hd_edit_desc &d, uint32_t x, uint32_t y, uint32_t z, bool &in_a,
                                                     bool &in_b) {
	const int32_t dx = int32_t(x - d.p0[0]), dy = int32_t(y - d.p0[1]), dz = int32_t(z - d.p0[2]); // world <= 2^22: no wrap
	const uint32_t ax = uint32_t(dx < 0 ? -dx : dx), ay = uint32_t(dy < 0 ? -dy : dy), az = uint32_t(dz < 0 ? -dz : dz) + 2u;
	if (max(ax, max(ay, az)) < 37837u) {
		const uint32_t s = uint32_t(dx * dx) + uint32_t(dy * dy);
		const uint32_t na = s + uint32_t(dz * dz), nb = s + uint32_t((dz + 2) * (dz + 2));
		in_a = uint64_t(na) <= d.r2, in_b = uint64_t(nb) <= d.r2;
	} else {
		const long long X = dx, Y = dy, Z = dz;
		const uint64_t s = uint64_t(X * X + Y * Y);
		in_a = s + uint64_t(Z * Z) <= d.r2, in_b = s + uint64_t((Z + 2) * (Z + 2)) <= d.r2;
	}
}

// Terrain fill of one 4x4x4 leaf, warp-cooperative (all 32 lanes call it with the same edit and leaf origin).
// A leaf's 4x4 footprint lies inside ONE lattice cell of every octave whose cell is >= 4 voxels wide, so the four
// lattice corners of each octave are hashed once per warp (lane 4*o + k hashes corner k of octave o) and handed round
// with shuffles; lane l < 16 then interpolates the height of column (x = l & 3, z = l >> 2) and the voxel owners read
// it with one more shuffle: 16 height evaluations per leaf instead of 64, and ~1/4 of the hashing in each.
// The integer arithmetic is terrain_height's, value for value.  Returns false when the octave layout does not fit
// (more than 8 octaves or a cell narrower than a leaf); the caller then evaluates per voxel.
__device__ __forceinline__ bool terrain_leaf_pair(const hd_edit_desc &d, uint32_t ox, uint32_t oz, uint32_t vx, uint32_t vy,
                                                  uint32_t vz, bool &in_a, bool &in_b) {
	const uint32_t lane = threadIdx.x & 31u, full = 0xFFFFFFFFu;
	uint32_t n_oct = 0;
	TerrainOctave oc;
	bool fits = true;
	for (; terrain_octave(d, n_oct, oc); ++n_oct)
		fits = fits && oc.cell_bits >= 2u;
	if (!fits || n_oct > 8u)
		return false;
	uint32_t lat = 0;
	if (terrain_octave(d, lane >> 2, oc))
		lat = terrain_lattice(oc.hseed, (ox >> oc.cell_bits) + (lane & 1u), (oz >> oc.cell_bits) + ((lane >> 1) & 1u));
	const uint32_t cx = ox + (lane & 3u), cz = oz + ((lane >> 2) & 3u); // this lane's column (lanes 16..31 repeat 0..15)
	uint32_t h = d.p0[0];
	for (uint32_t o = 0; o < n_oct; ++o) {
		const uint32_t v00 = __shfl_sync(full, lat, 4u * o), v10 = __shfl_sync(full, lat, 4u * o + 1u),
		               v01 = __shfl_sync(full, lat, 4u * o + 2u), v11 = __shfl_sync(full, lat, 4u * o + 3u);
		const uint32_t c = d.p0[1] - 2u * o, m = (1u << c) - 1u;
		const uint32_t r = terrain_bilerp(v00, v10, v01, v11, cx & m, cz & m, c);
		h += uint32_t((uint64_t(r) * (d.p1[0] >> (2u * o))) >> 16);
	}
	const uint32_t col = ((vz & 3u) << 2) | (vx & 3u);
	const uint32_t ha = __shfl_sync(full, h, col), hb = __shfl_sync(full, h, col + 8u); // voxel b: z + 2
	bool inside = true;
	if (d.p1[1] != 0u)
		inside = (vx >> d.p1[1]) == 0u && (vz >> d.p1[1]) == 0u; // z + 2 stays inside the same 4-aligned leaf
	in_a = inside && vy < ha;
	in_b = inside && vy < hb;
	return true;
}

// The same for the four voxels a lane owns when a HALF-warp works on a leaf: (x, y + 2*j, z + 2*k), j, k in {0, 1};
// in[j + 2*k].
__device__ __forceinline__ void sphere_in_range_quad(const hd_edit_desc &d, uint32_t x, uint32_t y, uint32_t z, bool (&in)[4]) {
	const int32_t dx = int32_t(x - d.p0[0]), dy = int32_t(y - d.p0[1]), dz = int32_t(z - d.p0[2]);
	const uint32_t ax = uint32_t(dx < 0 ? -dx : dx), ay = uint32_t(dy < 0 ? -dy : dy) + 2u, az = uint32_t(dz < 0 ? -dz : dz) + 2u;
	if (max(ax, max(ay, az)) < 37837u) {
		const uint32_t sx = uint32_t(dx * dx);
		const uint32_t sy0 = sx + uint32_t(dy * dy), sy1 = sx + uint32_t((dy + 2) * (dy + 2));
		const uint32_t z0 = uint32_t(dz * dz), z1 = uint32_t((dz + 2) * (dz + 2));
		in[0] = uint64_t(sy0 + z0) <= d.r2, in[1] = uint64_t(sy1 + z0) <= d.r2;
		in[2] = uint64_t(sy0 + z1) <= d.r2, in[3] = uint64_t(sy1 + z1) <= d.r2;
	} else {
		const long long X = dx, Y = dy, Z = dz;
		const uint64_t sx = uint64_t(X * X);
		const uint64_t sy0 = sx + uint64_t(Y * Y), sy1 = sx + uint64_t((Y + 2) * (Y + 2));
		const uint64_t z0 = uint64_t(Z * Z), z1 = uint64_t((Z + 2) * (Z + 2));
		in[0] = sy0 + z0 <= d.r2, in[1] = sy1 + z0 <= d.r2, in[2] = sy0 + z1 <= d.r2, in[3] = sy1 + z1 <= d.r2;
	}
}

// EditVoxel (main.cpp:60-63,133-142)
template <bool kTerrain = true>
__device__ __forceinline__ bool edit_voxel(const hd_edit_desc &d, uint32_t x, uint32_t y, uint32_t z, bool voxel) {
	const bool in = voxel_in_range<kTerrain>(d, x, y, z);
	return d.kind == HD_EDIT_SPHERE_DIG ? (voxel && !in) : (voxel || in);
}
// EditVoxel for the lane's two voxels (z and z + 2) of a leaf; spheres share dx^2 + dy^2 between them.
template <bool kTerrain>
__device__ __forceinline__ void edit_voxel_pair(const hd_edit_desc &d, uint32_t x, uint32_t y, uint32_t z, bool &a, bool &b) {
	const uint32_t kind = d.kind;
	if (kind == HD_EDIT_SPHERE_FILL || kind == HD_EDIT_SPHERE_DIG) {
		bool ia, ib;
		sphere_in_range_pair(d, x, y, z, ia, ib);
		a = kind == HD_EDIT_SPHERE_DIG ? (a && !ia) : (a || ia);
		b = kind == HD_EDIT_SPHERE_DIG ? (b && !ib) : (b || ib);
	} else {
		a = edit_voxel<kTerrain>(d, x, y, z, a);
		b = edit_voxel<kTerrain>(d, x, y, z + 2u, b);
	}
}

// EditVoxel for a lane's four voxels of a leaf (half-warp per leaf, no terrain): v[j + 2*k] = voxel (x, y + 2j, z + 2k).
__device__ __forceinline__ void edit_voxel_quad(const hd_edit_desc &d, uint32_t x, uint32_t y, uint32_t z, bool (&v)[4]) {
	const uint32_t kind = d.kind;
	if (kind == HD_EDIT_SPHERE_FILL || kind == HD_EDIT_SPHERE_DIG) {
		bool in[4];
		sphere_in_range_quad(d, x, y, z, in);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			v[i] = kind == HD_EDIT_SPHERE_DIG ? (v[i] && !in[i]) : (v[i] || in[i]);
	} else {
#pragma unroll
		for (int i = 0; i < 4; ++i)
			v[i] = edit_voxel<false>(d, x, y + 2u * (i & 1), z + (i & 2), v[i]);
	}
}

} // namespace hd
