// C++ host-layer test: the intents of the reference's (stale) test/test.cpp:117-209, written against
// include/hashdag_b200.hpp.  Exit code 0 = all checks passed.  Needs a CUDA device.
#include "hashdag_b200.hpp"

#include <cstdio>
#include <cstdlib>

using namespace hashdag_b200;

#define CHECK(cond)                                                                                                    \
	do {                                                                                                               \
		if (!(cond)) {                                                                                                 \
			std::fprintf(stderr, "CHECK failed: %s (%s:%d) last error: %s\n", #cond, __FILE__, __LINE__, hd_last_error()); \
			std::exit(1);                                                                                              \
		}                                                                                                              \
	} while (0)

struct SingleIterator { // test/test.cpp:36-52
	uint32_t level;
	UVec3 pos;
	bool exist;
	IterateType IterateNode(const NodeCoord<uint32_t> &coord, NodePointer<uint32_t> node) const {
		UVec3 lb = coord.GetLowerBoundAtLevel(level), ub = coord.GetUpperBoundAtLevel(level);
		bool overlaps = ub.x > pos.x && ub.y > pos.y && ub.z > pos.z && lb.x <= pos.x && lb.y <= pos.y && lb.z <= pos.z;
		return node && overlaps ? IterateType::kProceed : IterateType::kStop;
	}
	void IterateVoxel(const NodeCoord<uint32_t> &coord, bool voxel) {
		CHECK(coord.level == level);
		if (voxel && coord.pos == pos)
			exist = true;
	}
};

struct CountIterator {
	uint64_t voxels = 0, leaves = 0;
	IterateType IterateNode(const NodeCoord<uint32_t> &, NodePointer<uint32_t> node) {
		return node ? IterateType::kProceed : IterateType::kStop;
	}
	void IterateVoxel(const NodeCoord<uint32_t> &, bool voxel) { voxels += voxel; }
};

int main() {
	CHECK(hd_device_count() > 0);
	{ // Config / DefaultConfig (Config.hpp)
		auto cfg = DefaultConfig<uint32_t>{.level_count = 10, .top_level_count = 9}();
		CHECK(cfg.GetNodeLevels() == 9 && cfg.GetWordsPerPage() == 512 && cfg.GetWordsPerBucket() == 2048);
		CHECK(cfg.GetTotalBuckets() == 9216 && cfg.GetTotalWords() == 18874368ull && cfg.GetResolution() == 1024);
		auto bad = DefaultConfig<uint32_t>{.level_count = 17, .word_bits_per_page = 14, .bucket_bits_per_top_level = 10}();
		CHECK(!Config<uint32_t>::Validate(bad) && DAGNodePool::Create(bad) == nullptr);
	}
	{ // "Test upsert()" (test/test.cpp:118-147) through the C ABI
		auto pool = DAGNodePool::Create(DefaultConfig<uint32_t>{.level_count = 5}());
		CHECK(pool);
		uint32_t node0[3] = {0b11u, 0x2300, 0x4500}, p0, p1;
		CHECK(hd_upsert_nodes(pool->Handle(), 0, node0, 3, 1, &p0) == HD_OK && p0 != HD_NULL_NODE);
		CHECK(hd_upsert_nodes(pool->Handle(), 0, node0, 3, 1, &p1) == HD_OK && p1 == p0);
		uint32_t bw;
		CHECK(hd_pool_read_bucket_words(pool->Handle(), p0 / pool->GetConfig().GetWordsPerBucket(), &bw, 1) == HD_OK && bw == 3);
		uint32_t leaf[2] = {0x23, 0x55}, p6;
		CHECK(hd_upsert_nodes(pool->Handle(), 3, leaf, 2, 1, &p6) == HD_OK && p6 != HD_NULL_NODE);
		CHECK(hd_pool_read_bucket_words(pool->Handle(), p6 / pool->GetConfig().GetWordsPerBucket(), &bw, 1) == HD_OK && bw == 2);
	}
	{ // "Test Edit() and Iterate()" (test/test.cpp:161-199)
		auto pool = DAGNodePool::Create(DefaultConfig<uint32_t>{.level_count = 5}());
		CHECK(pool);
		const uint32_t vl = pool->GetConfig().GetVoxelLevel();
		auto root = pool->Edit({}, AABBEditor{{0, 0, 0}, {4, 4, 4}});
		CHECK(root);
		auto root2 = pool->Edit(root, AABBEditor{{0, 0, 0}, {4, 4, 4}});
		CHECK(root2 && root == root2);
		SingleIterator it{vl, {3, 3, 3}, false};
		pool->Iterate(root2, &it);
		CHECK(it.exist);
		it = SingleIterator{vl, {4, 3, 3}, false};
		pool->Iterate(root2, &it);
		CHECK(!it.exist);
		it = SingleIterator{vl, {3, 3, 3}, false};
		pool->Iterate({}, &it);
		CHECK(!it.exist);
		auto root3 = pool->Edit(root2, AABBEditor{{1, 1, 1}, {5, 5, 5}});
		CHECK(root3 && root != root3);
		auto root4 = pool->Edit(root3, AABBEditor{{1, 2, 3}, {3, 5, 5}});
		CHECK(root4 && root3 == root4);
		it = SingleIterator{vl, {4, 3, 3}, false};
		pool->Iterate(root4, &it);
		CHECK(it.exist);
		CountIterator cnt;
		pool->Iterate(root4, &cnt);
		CHECK(cnt.voxels == 64 + 64 - 27); // |[0,4)^3 u [1,5)^3|
	}
	{ // "Test ThreadedEdit()" (test/test.cpp:200-208) + sphere editors + pick ray (main.cpp:320-343)
		auto pool = DAGNodePool::Create(DefaultConfig<uint32_t>{.level_count = 10, .top_level_count = 9}());
		CHECK(pool);
		auto root = pool->ThreadedEdit(nullptr, {}, AABBEditor{{0, 0, 0}, {43, 21, 3}});
		CHECK(root);
		root = pool->ThreadedEdit(nullptr, {}, SphereEditor<EditMode::kFill>{{512, 512, 512}, 341ull * 341ull}, 10);
		CHECK(root && pool->GetLastEditStats().overflow_count == 0);
		pool->SetRoot(root);
		CHECK(pool->GetRoot() == root);
		auto hit = pool->Traversal<float>(root, {0.5f, 0.5f, 1.5f}, {0.f, 0.f, -1.f});
		CHECK(hit && hit->z > 0.8339f && hit->z < 0.8341f); // top of voxel 853 = 512+341
		CHECK(!pool->Traversal<float>(root, {0.5f, 0.5f, 1.5f}, {0.5145f, 0.f, -0.8575f}));
		CHECK(!pool->Traversal<float>({}, {0.5f, 0.5f, 1.5f}, {0.f, 0.f, -1.f}));
		auto dug = pool->Edit(root, SphereEditor<EditMode::kDig>{{512, 512, 853}, 40ull * 40ull});
		CHECK(pool->GetLastEditPath() == 1); // a single brush editor takes the one-launch low-latency path
		CHECK(dug && dug != root);
		auto hit2 = pool->Traversal<float>(dug, {0.5f, 0.5f, 1.5f}, {0.f, 0.f, -1.f});
		CHECK(hit2 && hit2->z < hit->z);
		CHECK(pool->Traversal<float>(root, {0.5f, 0.5f, 1.5f}, {0.f, 0.f, -1.f})->z == hit->z); // old root still valid
		// batched: 3 edits in one pass == 3 sequential edits (canonically: same voxel count here)
		hd_edit_desc batch[3] = {SphereEditor<EditMode::kFill>{{300, 300, 300}, 2500}.Desc(),
		                         SphereEditor<EditMode::kDig>{{310, 300, 300}, 900}.Desc(), AABBEditor{{280, 280, 280}, {290, 290, 290}}.Desc()};
		auto a = pool->EditBatch({}, batch, 3);
		auto b = pool->Edit(pool->Edit(pool->Edit({}, SphereEditor<EditMode::kFill>{{300, 300, 300}, 2500}),
		                               SphereEditor<EditMode::kDig>{{310, 300, 300}, 900}),
		                    AABBEditor{{280, 280, 280}, {290, 290, 290}});
		CHECK(a && a == b); // same pool: identical content dedups to the identical pointer
		// GC (main.cpp:235-238,373-378): the dug version survives, its pick ray is unchanged; save / load round trip
		auto kept = pool->ThreadedGC(nullptr, dug);
		CHECK(kept && pool->GetLastStatus() == HD_OK);
		auto hit3 = pool->Traversal<float>(kept, {0.5f, 0.5f, 1.5f}, {0.f, 0.f, -1.f});
		CHECK(hit3 && hit3->z == hit2->z);
		pool->SetRoot(kept);
		CHECK(pool->Save("/tmp/hashdag_b200_cpp_test.hdag"));
		auto loaded = DAGNodePool::Load("/tmp/hashdag_b200_cpp_test.hdag");
		CHECK(loaded && loaded->GetRoot() == kept && loaded->GetConfig().GetNodeLevels() == 9);
		auto hit4 = loaded->Traversal<float>(loaded->GetRoot(), {0.5f, 0.5f, 1.5f}, {0.f, 0.f, -1.f});
		CHECK(hit4 && hit4->z == hit2->z);
	}
	{ // colour-aware edits with the reference's wrapper names and its on_edit_done(root, state) contract (main.cpp:214-279)
		auto pool = DAGNodePool::Create(DefaultConfig<uint32_t>{.level_count = 8, .top_level_count = 9}());
		CHECK(pool && pool->ConfigureColor(4));
		struct EditResult {
			NodePointer<uint32_t> node_ptr;
			std::optional<uint32_t> opt_color_ptr;
		};
		const auto edit = [&](auto &&editor) -> EditResult {
			return pool->ThreadedEdit(nullptr, pool->GetRoot(), editor, 4, [&](NodePointer<uint32_t> root_ptr, auto &&state) -> EditResult {
				if constexpr (requires { state.octree_node; })
					return {root_ptr, state.octree_node};
				else
					return {root_ptr, std::nullopt};
			});
		};
		auto r1 = edit(VBREditorWrapper<AABBEditor>{{{10, 10, 10}, {100, 60, 100}, 0xFFFFFF}});
		CHECK(r1.node_ptr && r1.opt_color_ptr && *r1.opt_color_ptr == pool->GetColorRoot() && pool->GetLastStatus() == HD_OK);
		pool->SetRoot(r1.node_ptr);
		auto r2 = edit(VBREditorWrapper<SphereEditor<EditMode::kPaint>>{{{50, 40, 50}, 900, 0x007FFF}});
		CHECK(r2.node_ptr == r1.node_ptr);                                        // paint leaves the geometry alone
		CHECK(r2.opt_color_ptr && *r2.opt_color_ptr != *r1.opt_color_ptr);        // ... and changes the colour octree
		auto r3 = edit(StatelessEditorWrapper<SphereEditor<EditMode::kDig>>{{{50, 60, 50}, 400}});
		CHECK(!r3.opt_color_ptr && r3.node_ptr && r3.node_ptr != r2.node_ptr);    // stateless: no colour state
		auto r4 = edit(SphereEditor<EditMode::kFill>{{30, 30, 30}, 100});          // a bare editor is stateless too
		CHECK(!r4.opt_color_ptr && r4.node_ptr);
		// Iterate over a tree of a few thousand nodes: whole subtrees come back per device round trip
		CountIterator cnt;
		const uint64_t launches0 = hd_kernel_launches();
		pool->Iterate(r1.node_ptr, &cnt);
		CHECK(cnt.voxels == 90ull * 50 * 90);
		CHECK(hd_kernel_launches() - launches0 < 2000);                           // far fewer read-backs than nodes
	}
	std::puts("cpp host api: OK");
	return 0;
}
