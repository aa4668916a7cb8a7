// tools/simt_model.cpp — offline SIMT cost model of the trace kernel's per-ray state machine (no GPU needed).
//
// The trace kernel is issue-bound with 19 of 32 lanes active per instruction (profiles/r1b_trace_kernel_ncu_full.txt),
// so what matters is how many warp instructions a scheduling of the PUSH / ADVANCE / POP state machine
// (shader/src/trace.frag:128-221) issues for a warp of 32 rays.  This tool replays real rays of a real scene through
// that state machine lane by lane, groups them in 8x4 pixel warps exactly like trace.cu, and charges every control-flow
// region the warp enters with the region's SASS instruction count.  It is a design tool for choosing between loop
// organisations before spending GPU time; numbers it prints are estimates, never bench values.
//
// Build: g++ -O2 -shared -fPIC -o tools/_simt_model.so tools/simt_model.cpp      (driven by tools/simt_model.py)
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../include/hashdag_b200.h"

namespace {
constexpr uint32_t kStack = 23;
inline float fbits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t ubits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float fmin2(float a, float b) { return b < a ? b : a; }
inline float fmax2(float a, float b) { return a < b ? b : a; }

enum Want { kFetch, kTest, kPush, kPop, kDone };
// per-scale census of the transitions (simt_hist): [scale][0 push, 1 advance, 2 pop (arrival scale)]
uint64_t g_hist[kStack + 1][3];

struct Lane {
	const uint32_t *nodes;
	float tc[3], tb[3], pos[3], t_min, t_max, h, scale_exp2;
	uint32_t stack[kStack], parent, child_bits, idx, scale, leaf_scale, octant;
	float t_corner[3], tc_max;
	uint32_t child_shift;
	bool done = true, hit = false;
	uint32_t ops = 0; // thread-level useful region executions (for the efficiency figure)

	void start(const uint32_t *n, uint32_t root, uint32_t leaf_level, const float o_in[3], const float d_in[3]) {
		nodes = n;
		const float eps = fbits((127u - kStack) << 23);
		float o[3], d[3];
		for (int i = 0; i < 3; ++i) {
			o[i] = o_in[i] + 1.0f;
			float dd = d_in[i];
			d[i] = std::fabs(dd) > eps ? dd : (dd >= 0 ? eps : -eps);
			tc[i] = 1.0f / -std::fabs(d[i]);
			tb[i] = tc[i] * o[i];
		}
		octant = 0;
		for (int i = 0; i < 3; ++i)
			if (d[i] > 0.0f)
				octant ^= 1u << i, tb[i] = 3.0f * tc[i] - tb[i];
		t_min = fmax2(fmax2(2.0f * tc[0] - tb[0], 2.0f * tc[1] - tb[1]), 2.0f * tc[2] - tb[2]);
		t_max = fmin2(fmin2(tc[0] - tb[0], tc[1] - tb[1]), tc[2] - tb[2]);
		h = t_max;
		t_min = fmax2(t_min, 0.0f);
		t_max = fmin2(t_max, 1.0f);
		parent = root, child_bits = 0, idx = 0;
		for (int i = 0; i < 3; ++i) {
			pos[i] = 1.0f;
			if (1.5f * tc[i] - tb[i] > t_min)
				idx ^= 1u << i, pos[i] = 1.5f;
		}
		scale = kStack - 1, scale_exp2 = 0.5f, leaf_scale = kStack - leaf_level;
		done = false, hit = false;
	}
	// 0: nothing to fetch (sub-leaf), 1: inner mask, 2: leaf pair
	int fetch_kind() const { return child_bits != 0 ? -1 : scale > leaf_scale ? 1 : scale == leaf_scale ? 2 : 0; }
	void fetch() {
		if (scale > leaf_scale)
			child_bits = nodes[parent];
		else if (scale == leaf_scale) {
			uint32_t l0 = nodes[parent], l1 = nodes[parent + 1];
			for (uint32_t b = 0; b < 4; ++b) {
				child_bits |= ((l0 >> (8 * b)) & 0xFFu) ? (1u << b) : 0u;
				child_bits |= ((l1 >> (8 * b)) & 0xFFu) ? (16u << b) : 0u;
			}
		} else
			child_bits = parent;
	}
	bool test() { // true: the current child exists and is in range (-> PUSH), false -> ADVANCE
		for (int i = 0; i < 3; ++i)
			t_corner[i] = pos[i] * tc[i] - tb[i];
		tc_max = fmin2(fmin2(t_corner[0], t_corner[1]), t_corner[2]);
		child_shift = idx ^ octant;
		return (child_bits & (1u << child_shift)) != 0 && t_min <= t_max;
	}
	void push() { // full detail: no LOD cut-off
		const uint32_t child_mask = 1u << child_shift;
		float half = scale_exp2 * 0.5f;
		float t_center[3] = {half * tc[0] + t_corner[0], half * tc[1] + t_corner[1], half * tc[2] + t_corner[2]};
		if (scale < leaf_scale) {
			done = true, hit = true;
			return;
		}
		if (tc_max < h)
			stack[scale] = parent;
		h = tc_max;
		g_hist[scale][0]++;
		if (scale > leaf_scale)
			parent = nodes[parent + 1u + uint32_t(__builtin_popcount(child_bits & (child_mask - 1u)))];
		else
			parent = (nodes[parent + (child_shift >> 2)] >> ((child_shift & 3u) << 3)) & 0xFFu;
		idx = 0, --scale, scale_exp2 = half;
		for (int i = 0; i < 3; ++i)
			if (t_center[i] > t_min)
				idx ^= 1u << i, pos[i] += scale_exp2;
		child_bits = 0;
	}
	uint32_t step_mask;
	bool advance() { // true: a POP must follow
		step_mask = 0;
		for (int i = 0; i < 3; ++i)
			if (t_corner[i] <= tc_max)
				step_mask ^= 1u << i, pos[i] -= scale_exp2;
		t_min = tc_max;
		idx ^= step_mask;
		g_hist[scale][1]++;
		return (idx & step_mask) != 0;
	}
	void pop() {
		uint32_t differing = 0;
		for (int i = 0; i < 3; ++i)
			if (step_mask >> i & 1u)
				differing |= ubits(pos[i]) ^ ubits(pos[i] + scale_exp2);
		scale = differing ? 31u - uint32_t(__builtin_clz(differing)) : 0xFFFFFFFFu;
		if (scale >= kStack) {
			done = true, hit = false;
			return;
		}
		scale_exp2 = fbits((scale - kStack + 127u) << 23);
		g_hist[scale][2]++;
		parent = stack[scale];
		uint32_t sh[3];
		for (int i = 0; i < 3; ++i)
			sh[i] = ubits(pos[i]) >> scale, pos[i] = fbits(sh[i] << scale);
		idx = (sh[0] & 1u) | ((sh[1] & 1u) << 1) | ((sh[2] & 1u) << 2);
		h = 0.0f;
		child_bits = 0;
	}
};

// instruction counts of the regions (SASS of trace_kernel<false,false,true>, round 1)
struct Costs {
	uint32_t loop = 4, fetch_inner = 5, fetch_leaf = 15, fetch_sub = 2, test = 14, push = 42, adv = 13, pop = 33, setup = 150,
	         inner_loop = 3, refill_check = 6;
};

struct Ray {
	float o[3], d[3];
};

enum Region { rLoop, rFetchInner, rFetchLeaf, rFetchSub, rTest, rPush, rAdv, rPop, rSetup, rCount };
struct Acc {
	uint64_t warp_instr = 0, thread_instr = 0, rays = 0, trips = 0, hits = 0;
	uint64_t execs[rCount] = {}, lanes_sum[rCount] = {}, cost_sum[rCount] = {};
	void region(uint32_t cost, uint32_t lanes, int r) {
		warp_instr += cost, thread_instr += uint64_t(cost) * lanes;
		execs[r]++, lanes_sum[r] += lanes, cost_sum[r] += cost;
	}
};

void make_ray(const hd_trace_params &P, uint32_t px, uint32_t py, Ray &r) {
	float cx = (float(px) + 0.5f) / float(P.width), cy = (float(py) + 0.5f) / float(P.height);
	cx = cx * 2.0f - 1.0f, cy = cy * 2.0f - 1.0f;
	for (int i = 0; i < 3; ++i)
		r.d[i] = (P.look[i] - P.side[i] * cx) - P.up[i] * cy, r.o[i] = P.pos[i];
	float dot = (r.d[0] * r.d[0] + r.d[1] * r.d[1]) + r.d[2] * r.d[2];
	float inv = 1.0f / std::sqrt(dot);
	for (int i = 0; i < 3; ++i)
		r.d[i] *= inv;
}

// One warp; `queue` = pixels it will process (first 32 start in the lanes, the rest refill when policy allows).
// policy bit 0: two-phase (advance runs until every lane wants PUSH / POP / is done); bit 1: refill idle lanes when at
// least `refill_min` lanes are idle; bit 2: POP deferred = lanes that want POP wait until no lane wants PUSH (and v.v.)
void run_warp(const uint32_t *nodes, const hd_trace_params &P, const std::vector<std::pair<uint32_t, uint32_t>> &queue, int policy,
              uint32_t refill_min, const Costs &C, Acc &acc) {
	Lane L[32];
	size_t next = 0;
	auto refill = [&](bool initial) {
		uint32_t n = 0;
		for (int l = 0; l < 32 && next < queue.size(); ++l)
			if (L[l].done) {
				Ray r;
				make_ray(P, queue[next].first, queue[next].second, r);
				++next;
				L[l].start(nodes, P.dag_root, P.dag_leaf_level, r.o, r.d);
				++n, ++acc.rays;
			}
		if (n)
			acc.region(C.setup, n, rSetup);
		(void)initial;
	};
	refill(true);
	const bool two_phase = policy & 1, do_refill = policy & 2;
	bool pend[32] = {};
	for (;;) {
		uint32_t active = 0;
		for (auto &l : L)
			active += !l.done;
		if (do_refill && next < queue.size() && 32 - active >= refill_min) {
			refill(false);
			active = 0;
			for (auto &l : L)
				active += !l.done;
		}
		if (!active) {
			if (next < queue.size()) {
				refill(false);
				continue;
			}
			break;
		}
		++acc.trips;
		acc.region(C.loop + (do_refill ? C.refill_check : 0), active, rLoop);
		// fetch phase
		uint32_t nf[3] = {0, 0, 0};
		for (int i = 0; i < 32; ++i)
			if (!L[i].done && !pend[i]) {
				int k = L[i].fetch_kind();
				if (k >= 0)
					++nf[k], L[i].fetch();
			}
		if (nf[1])
			acc.region(C.fetch_inner, nf[1], rFetchInner);
		if (nf[2])
			acc.region(C.fetch_leaf, nf[2], rFetchLeaf);
		if (nf[0])
			acc.region(C.fetch_sub, nf[0], rFetchSub);
		if (!two_phase) {
			const bool defer = policy & 8;
			uint32_t n_push = 0, n_adv = 0, n_pop = 0, n_wait = 0, n_test = 0;
			for (int i = 0; i < 32; ++i)
				if (!L[i].done) {
					if (pend[i]) {
						++n_wait;
						continue;
					}
					++n_test;
					if (L[i].test())
						++n_push, L[i].push();
					else {
						++n_adv;
						if (L[i].advance()) {
							if (defer)
								pend[i] = true, ++n_wait;
							else
								++n_pop, L[i].pop();
						}
					}
					if (L[i].done && L[i].hit)
						++acc.hits;
				}
			if (n_test)
				acc.region(C.test, n_test, rTest);
			if (n_push)
				acc.region(C.push, n_push, rPush);
			if (n_adv)
				acc.region(C.adv, n_adv, rAdv);
			if (defer && n_wait) {
				acc.region(3, active, rLoop); // ballot + popc + compare
				uint32_t still = 0;
				for (int i = 0; i < 32; ++i)
					still += !L[i].done && !pend[i];
				if (n_wait >= refill_min || still == 0) {
					for (int i = 0; i < 32; ++i)
						if (pend[i])
							pend[i] = false, L[i].pop();
					n_pop = n_wait;
				}
			}
			if (n_pop)
				acc.region(C.pop, n_pop, rPop);
		} else {
			// inner loop: every lane advances until it wants PUSH or POP
			Want w[32];
			for (int i = 0; i < 32; ++i)
				w[i] = L[i].done ? kDone : kTest;
			for (;;) {
				uint32_t n_test = 0;
				for (int i = 0; i < 32; ++i)
					n_test += w[i] == kTest;
				if (!n_test)
					break;
				acc.region(C.test + C.inner_loop, n_test, rTest);
				uint32_t n_adv = 0;
				for (int i = 0; i < 32; ++i)
					if (w[i] == kTest) {
						if (L[i].test())
							w[i] = kPush;
						else {
							++n_adv;
							if (L[i].advance())
								w[i] = kPop;
						}
					}
				if (n_adv)
					acc.region(C.adv, n_adv, rAdv);
			}
			uint32_t n_push = 0, n_pop = 0;
			for (int i = 0; i < 32; ++i) {
				if (w[i] == kPush)
					++n_push, L[i].push();
				else if (w[i] == kPop)
					++n_pop, L[i].pop();
				if (w[i] != kDone && L[i].done && L[i].hit)
					++acc.hits;
			}
			if (n_push)
				acc.region(C.push, n_push, rPush);
			if (n_pop)
				acc.region(C.pop, n_pop, rPop);
		}
	}
}
// Greedy region scheduling: every lane is in one of {wants TEST(+fetch, +advance), wants PUSH, wants POP, done}; each step
// the warp executes the ONE region most lanes wait for (a ballot + popc decision, charged C.refill_check instructions).
// `bias` > 0 favours the TEST region (it is the cheapest and feeds the other two).
void run_warp_greedy(const uint32_t *nodes, const hd_trace_params &P, const std::vector<std::pair<uint32_t, uint32_t>> &queue,
                     uint32_t pop_min, const Costs &C, Acc &acc) {
	Lane L[32];
	Want w[32];
	uint32_t n = 0;
	for (size_t i = 0; i < 32 && i < queue.size(); ++i) {
		Ray r;
		make_ray(P, queue[i].first, queue[i].second, r);
		L[i].start(nodes, P.dag_root, P.dag_leaf_level, r.o, r.d);
		++n, ++acc.rays;
	}
	acc.region(C.setup, n, rSetup);
	for (int i = 0; i < 32; ++i)
		w[i] = L[i].done ? kDone : kTest;
	for (;;) {
		uint32_t nt = 0, np = 0, no = 0;
		for (int i = 0; i < 32; ++i)
			nt += w[i] == kTest, np += w[i] == kPush, no += w[i] == kPop;
		if (!nt && !np && !no)
			break;
		++acc.trips;
		acc.region(C.loop + C.refill_check, nt + np + no, rLoop);
		// choice: POP only when at least pop_min lanes want it or nothing else can run; otherwise the larger of TEST / PUSH
		int choice;
		if (no && (no >= pop_min || (!nt && !np)))
			choice = no >= std::max(nt, np) || (!nt && !np) ? 2 : (nt >= np ? 0 : 1);
		else
			choice = nt >= np && nt ? 0 : (np ? 1 : 0);
		if (choice == 0 && !nt)
			choice = np ? 1 : 2;
		if (choice == 0) {
			uint32_t nf[3] = {0, 0, 0};
			for (int i = 0; i < 32; ++i)
				if (w[i] == kTest) {
					int k = L[i].fetch_kind();
					if (k >= 0)
						++nf[k], L[i].fetch();
				}
			if (nf[1])
				acc.region(C.fetch_inner, nf[1], rFetchInner);
			if (nf[2])
				acc.region(C.fetch_leaf, nf[2], rFetchLeaf);
			if (nf[0])
				acc.region(C.fetch_sub, nf[0], rFetchSub);
			acc.region(C.test, nt, rTest);
			uint32_t n_adv = 0;
			for (int i = 0; i < 32; ++i)
				if (w[i] == kTest) {
					if (L[i].test())
						w[i] = kPush;
					else {
						++n_adv;
						if (L[i].advance())
							w[i] = kPop;
					}
				}
			if (n_adv)
				acc.region(C.adv, n_adv, rAdv);
		} else if (choice == 1) {
			for (int i = 0; i < 32; ++i)
				if (w[i] == kPush) {
					L[i].push();
					w[i] = L[i].done ? kDone : kTest;
					acc.hits += L[i].done && L[i].hit;
				}
			acc.region(C.push, np, rPush);
		} else {
			for (int i = 0; i < 32; ++i)
				if (w[i] == kPop) {
					L[i].pop();
					w[i] = L[i].done ? kDone : kTest;
				}
			acc.region(C.pop, no, rPop);
		}
	}
}
} // namespace

extern "C" {
// per-scale transition census accumulated by every simt_model call so far; clear = 1 resets it afterwards
void simt_hist(uint64_t *out /* [24][3] */, int clear) {
	std::memcpy(out, g_hist, sizeof(g_hist));
	if (clear)
		std::memset(g_hist, 0, sizeof(g_hist));
}
// out[0..4] = warp instructions, thread instructions, rays, warp trips, hits; then per region {executions, lanes, cost}.
// Samples every `cta_step`-th 16x8 CTA patch of the frame.  `per_warp_patches` > 1 gives each warp that many 8x4 patches
// (stacked vertically) as its refill queue.
void simt_model(const uint32_t *nodes, const hd_trace_params *P, int policy, uint32_t refill_min, uint32_t per_warp_patches,
                uint32_t cta_step, const uint32_t *costs /* 11 values or NULL */, uint64_t *out) {
	Costs C;
	if (costs) {
		C.loop = costs[0], C.fetch_inner = costs[1], C.fetch_leaf = costs[2], C.fetch_sub = costs[3], C.test = costs[4];
		C.push = costs[5], C.adv = costs[6], C.pop = costs[7], C.setup = costs[8], C.inner_loop = costs[9], C.refill_check = costs[10];
	}
	Acc acc;
	const uint32_t W = P->width, H = P->height;
	const uint32_t strips_y = (H + 4 * per_warp_patches - 1) / (4 * per_warp_patches), strips_x = (W + 7) / 8;
	uint32_t k = 0;
	for (uint32_t sy = 0; sy < strips_y; ++sy)
		for (uint32_t sx = 0; sx < strips_x; ++sx, ++k) {
			if (k % cta_step)
				continue;
			std::vector<std::pair<uint32_t, uint32_t>> q;
			for (uint32_t p = 0; p < per_warp_patches; ++p)
				for (uint32_t l = 0; l < 32; ++l) {
					uint32_t px = sx * 8 + (l & 7), py = (sy * per_warp_patches + p) * 4 + (l >> 3);
					if (px < W && py < H)
						q.emplace_back(px, py);
				}
			if (policy & 4)
				run_warp_greedy(nodes, *P, q, refill_min, C, acc);
			else
				run_warp(nodes, *P, q, policy, refill_min, C, acc);
		}
	out[0] = acc.warp_instr, out[1] = acc.thread_instr, out[2] = acc.rays, out[3] = acc.trips, out[4] = acc.hits;
	for (int r = 0; r < rCount; ++r)
		out[5 + 3 * r] = acc.execs[r], out[6 + 3 * r] = acc.lanes_sum[r], out[7 + 3 * r] = acc.cost_sum[r];
}
}
