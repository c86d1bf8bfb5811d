// CPU test of hcm_host::ShardedMeshDispatcher (hypercraft_b200/host/hcm_host.hpp) against a RECORDING FAKE
// of the C ABI with one fake ctx per GPU -- test infrastructure only, linked instead of libhcm.so.  The fake's hcm_set_blocks
// reports the reference's remesh fan-out the way the device does: only positions whose 27 chunks lie inside that ctx's
// window.  Checked: tasks run on the slab owner, each ctx only ever sees chunks of its own window, the boundary planes go to
// both neighbours, an edit next to a cut is applied by the owner alone, remeshes the tasks across the cut as well, and makes
// the other GPU take the edited chunk again.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <map>
#include <set>

#include "../../hypercraft_b200/host/hcm_host.hpp"

using P3 = std::array<int, 3>;
extern "C" {
struct hcm_ctx {
	hcm_config cfg;
	std::vector<std::vector<P3>> uploads, tasks; // per hcm_upload_and_mesh call
	std::vector<hcm_block_edit> edits;
	std::vector<hcm_sunlight_change> sun;
	std::vector<int16_t> remesh;
	std::vector<uint8_t> remesh_flags;
	std::vector<hcm_task_result> results;
	std::set<P3> resident;
	std::vector<std::pair<const hcm_ctx *, std::vector<P3>>> peer_copies; // hcm_copy_chunks_from calls received: (source, positions)
	bool in_window(int x, int y, int z) const {
		return x >= cfg.origin[0] && x < cfg.origin[0] + cfg.extent[0] && y >= cfg.origin[1] && y < cfg.origin[1] + cfg.extent[1] && z >= cfg.origin[2] &&
		       z < cfg.origin[2] + cfg.extent[2];
	}
};
}
static std::vector<hcm_ctx *> g_ctx;
static std::atomic<int> g_violations{0};

extern "C" {
int hcm_registry_from_blob(const void *, size_t, hcm_registry *out) { return *out = hcm_registry{}, HCM_OK; }
int hcm_create(const hcm_registry *, const hcm_config *cfg, hcm_ctx **out) {
	*out = new hcm_ctx{};
	(*out)->cfg = *cfg;
	g_ctx.push_back(*out);
	return HCM_OK;
}
void hcm_destroy(hcm_ctx *) {} // kept for inspection
const char *hcm_last_error(const hcm_ctx *) { return "fake"; }
int hcm_remove_chunks(hcm_ctx *c, uint32_t n, const int16_t *pos) {
	for (uint32_t i = 0; i < n; ++i)
		c->resident.erase({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
	return HCM_OK;
}
int hcm_is_resident(const hcm_ctx *c, const int16_t pos[3]) { return (int)c->resident.count({pos[0], pos[1], pos[2]}); }
int hcm_copy_chunks_from(hcm_ctx *dst, hcm_ctx *src, uint32_t n, const int16_t *pos) {
	dst->peer_copies.emplace_back(src, std::vector<P3>{});
	for (uint32_t i = 0; i < n; ++i) {
		const P3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
		g_violations += !dst->in_window(p[0], p[1], p[2]) || !src->resident.count(p);
		dst->peer_copies.back().second.push_back(p);
		dst->resident.insert(p);
	}
	return HCM_OK;
}
int hcm_upload_and_mesh(hcm_ctx *c, uint32_t n_chunks, const int16_t *pos, const uint16_t *, const int8_t *, uint32_t n_tasks, const int16_t *tp,
                        hcm_batch_view *out) {
	c->uploads.emplace_back(), c->tasks.emplace_back();
	for (uint32_t i = 0; i < n_chunks; ++i) {
		g_violations += !c->in_window(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		c->uploads.back().push_back({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
		c->resident.insert({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
	}
	for (uint32_t t = 0; t < n_tasks; ++t) {
		g_violations += !c->in_window(tp[3 * t] - 1, tp[3 * t + 1] - 1, tp[3 * t + 2] - 1) || !c->in_window(tp[3 * t] + 1, tp[3 * t + 1] + 1, tp[3 * t + 2] + 1);
		c->tasks.back().push_back({tp[3 * t], tp[3 * t + 1], tp[3 * t + 2]});
	}
	c->results.assign(n_tasks, hcm_task_result{0, 0}); // zero meshes is a valid result and is pushed too
	*out = hcm_batch_view{n_tasks, 0, 0, 0, c->results.data(), nullptr, nullptr, nullptr};
	return HCM_OK;
}
// every edit changes its block; a stone block (1) also raises the sunlight height of its column in its chunk
int hcm_set_blocks(hcm_ctx *c, uint32_t n, const hcm_block_edit *e, hcm_edit_result *out) {
	c->sun.clear(), c->remesh.clear();
	std::set<P3> fan;
	auto report = [&](int x, int y, int z) { // the device reports a position only if it sees all 27 chunks
		if (c->in_window(x - 1, y - 1, z - 1) && c->in_window(x + 1, y + 1, z + 1))
			fan.insert({x, y, z});
	};
	for (uint32_t i = 0; i < n; ++i) {
		g_violations += !c->in_window(e[i].chunk[0], e[i].chunk[1], e[i].chunk[2]);
		c->edits.push_back(e[i]);
		const int bx = e[i].index & 31, by = e[i].index >> 10, bz = (e[i].index >> 5) & 31;
		for (int dx = -1; dx <= 1; ++dx)
			for (int dy = -1; dy <= 1; ++dy)
				for (int dz = -1; dz <= 1; ++dz)
					if ((!dx && !dy && !dz) || hcm_host::ShardedMeshDispatcher<int>::ShouldRemesh(bx - 32 * dx, by - 32 * dy, bz - 32 * dz))
						report(e[i].chunk[0] + dx, e[i].chunk[1] + dy, e[i].chunk[2] + dz);
		if (e[i].block == 1) {
			c->sun.push_back(hcm_sunlight_change{{e[i].chunk[0], e[i].chunk[1], e[i].chunk[2]}, (uint16_t)(bx + bz * 32), 0, (int8_t)(by + 1), 0});
			for (int dx = -1; dx <= 1; ++dx)
				for (int dy = -1; dy <= 1; ++dy)
					for (int dz = -1; dz <= 1; ++dz)
						report(e[i].chunk[0] + dx, e[i].chunk[1] + dy, e[i].chunk[2] + dz);
		}
	}
	for (const P3 &p : fan)
		for (int k = 0; k < 3; ++k)
			c->remesh.push_back((int16_t)p[k]);
	c->remesh_flags.assign(fan.size(), (uint8_t)(c->sun.empty() ? HCM_REMESH_BLOCK_EDIT : HCM_REMESH_BLOCK_EDIT | HCM_REMESH_SUNLIGHT));
	*out = hcm_edit_result{n, (uint32_t)c->sun.size(), (uint32_t)fan.size(), 0, c->sun.data(), c->remesh.data(), c->remesh_flags.data()};
	return HCM_OK;
}
}

struct Pos {
	int x, y, z;
};
struct Chunk {
	Pos pos;
	std::vector<uint16_t> blocks = std::vector<uint16_t>(HCM_CHUNK_VOXELS);
	std::vector<int8_t> heights = std::vector<int8_t>(HCM_CHUNK_COLUMNS);
	const uint16_t *GetBlockData() const { return blocks.data(); }
	int8_t GetSunlightHeight(uint32_t i) const { return heights[i]; }
	const Pos &GetPosition() const { return pos; }
};

static int fails = 0;
#define CHECK(c)                                                                                                       \
	do {                                                                                                               \
		if (!(c)) {                                                                                                    \
			printf("FAILED line %d: %s\n", __LINE__, #c);                                                             \
			++fails;                                                                                                   \
		}                                                                                                              \
	} while (0)

int main() {
	using hcm_host::ChunkPos3;
	using D = hcm_host::ShardedMeshDispatcher<Chunk>;
	std::map<P3, std::shared_ptr<Chunk>> world; // chunks x in [-4,3], y and z in [-2,2]; tasks x in [-3,2], y and z in [-1,1]
	for (int x = -4; x <= 3; ++x)
		for (int y = -2; y <= 2; ++y)
			for (int z = -2; z <= 2; ++z)
				world[{x, y, z}] = std::make_shared<Chunk>(Chunk{{x, y, z}});
	auto gather = [&](int x, int y, int z) {
		std::array<std::shared_ptr<Chunk>, 27> a;
		int i = 0;
		for (int dx = -1; dx <= 1; ++dx)
			for (int dy = -1; dy <= 1; ++dy)
				for (int dz = -1; dz <= 1; ++dz)
					if (dx || dy || dz)
						a[i++] = world.at({x + dx, y + dy, z + dz});
		a[26] = world.at({x, y, z});
		return a;
	};
	auto is_task = [](const ChunkPos3 &p) { return p.x >= -3 && p.x <= 2 && p.y >= -1 && p.y <= 1 && p.z >= -1 && p.z <= 1; };
	const char blob[4] = {};
	bool threw = false;
	try {
		D bad(blob, sizeof blob, {{0, -3, 0}, {1, 1, 3}}, -2, -2, 5, 5); // a gap between the slabs
	} catch (const hcm_host::Error &) {
		threw = true;
	}
	CHECK(threw);
	for (const bool peer : {true, false}) { // halo replication device to device (the default), and through the host
	g_ctx.clear();
	g_violations = 0;
	for (auto &kv : world) // fresh chunk objects: nothing is resident anywhere
		kv.second = std::make_shared<Chunk>(Chunk{{kv.first[0], kv.first[1], kv.first[2]}});
	D disp(blob, sizeof blob, {{0, -3, 0}, {1, 0, 3}}, -2, -2, 5, 5, peer);
	CHECK(disp.Ranks() == 2 && g_ctx.size() == 2);
	CHECK(g_ctx[0]->cfg.device == 0 && g_ctx[0]->cfg.origin[0] == -4 && g_ctx[0]->cfg.extent[0] == 5); // planes -4 .. 0
	CHECK(g_ctx[1]->cfg.device == 1 && g_ctx[1]->cfg.origin[0] == -1 && g_ctx[1]->cfg.extent[0] == 5); // planes -1 .. 3
	CHECK(disp.Owner(-3) == 0 && disp.Owner(-1) == 0 && disp.Owner(0) == 1 && disp.Owner(2) == 1 && disp.Owner(3) == -1 && disp.Owner(-4) == -1);
	CHECK(disp.Holders(-2) == std::vector<int>{0} && disp.Holders(-1) == (std::vector<int>{0, 1}) && disp.Holders(0) == (std::vector<int>{0, 1}) &&
	      disp.Holders(1) == std::vector<int>{1});

	std::vector<ChunkPos3> sunk;
	auto sink = [&](const ChunkPos3 &p, std::vector<hcm_host::BlockMesh> &&m) {
		CHECK(m.empty());
		sunk.push_back(p);
	};
	// world load: every task once
	for (int x = -3; x <= 2; ++x)
		for (int y = -1; y <= 1; ++y)
			for (int z = -1; z <= 1; ++z)
				disp.Submit(gather(x, y, z));
	CHECK(disp.Pending() == 54);
	disp.Flush(sink);
	CHECK(sunk.size() == 54 && disp.Pending() == 0 && g_violations == 0);
	for (int r = 0; r < 2; ++r) {
		CHECK(g_ctx[r]->tasks.size() == 1 && g_ctx[r]->tasks[0].size() == 27 && g_ctx[r]->uploads[0].size() == 5 * 25);
		for (const P3 &t : g_ctx[r]->tasks[0])
			CHECK(disp.Owner(t[0]) == r);
	}
	auto has = [](const std::vector<P3> &v, P3 p) { return std::find(v.begin(), v.end(), p) != v.end(); };
	CHECK(has(g_ctx[0]->uploads[0], {0, 0, 0}) && has(g_ctx[1]->uploads[0], {0, 0, 0}) && has(g_ctx[0]->uploads[0], {-1, 1, 1}) && has(g_ctx[1]->uploads[0], {-1, 1, 1}));

	// an edit on the last voxel plane before the cut, owned by GPU 0: glowstone at (31, 16, 16) of chunk (-1,0,0)
	sunk.clear();
	std::vector<D::SetBlock> edits{{{-1, 0, 0}, (uint16_t)(31 + (16 * 32 + 16) * 32), 24}};
	world.at({-1, 0, 0})->blocks[edits[0].index] = 24; // the host's chunk is the authority (Chunk::SetBlock)
	std::vector<ChunkPos3> asked;
	int n_sun = 0;
	uint32_t changed = disp.SetBlocks(
	    edits, [&](const D::SunlightChange &) { ++n_sun; },
	    [&](const ChunkPos3 &p, bool) {
		    asked.push_back(p);
		    if (is_task(p))
			    disp.Submit(gather(p.x, p.y, p.z), true);
	    });
	CHECK(changed == 1 && n_sun == 0 && g_ctx[0]->edits.size() == 1 && g_ctx[1]->edits.empty());
	auto asked_has = [&](ChunkPos3 p) { return std::find(asked.begin(), asked.end(), p) != asked.end(); };
	CHECK(asked_has({-1, 0, 0}) && asked_has({0, 0, 0}));   // the chunk itself (reported by GPU 0) and the task across the cut (derived)
	CHECK(!asked_has({-2, 0, 0}) && !asked_has({1, 0, 0})); // 31 voxels from the far face, 33 from the next chunk over: no remesh
	{
		std::set<P3> unique;
		for (const ChunkPos3 &p : asked)
			unique.insert({p.x, p.y, p.z});
		CHECK(unique.size() == asked.size()); // each position asked once
	}
	disp.Flush(sink);
	CHECK(g_violations == 0 && g_ctx[0]->tasks.size() == 2 && g_ctx[1]->tasks.size() == 2);
	CHECK(g_ctx[0]->uploads[1].empty());                                                    // the owner edited on the device
	if (peer) { // the halo holder received the chunk from the owner's GPU, nothing came from the host
		CHECK(g_ctx[1]->uploads[1].empty() && g_ctx[1]->peer_copies.size() == 1 && g_ctx[1]->peer_copies[0].first == g_ctx[0] &&
		      (g_ctx[1]->peer_copies[0].second == std::vector<P3>{P3{-1, 0, 0}}) && g_ctx[0]->peer_copies.empty());
	} else
		CHECK(g_ctx[1]->uploads[1].size() == 1 && g_ctx[1]->uploads[1][0] == (P3{-1, 0, 0}) && g_ctx[1]->peer_copies.empty()); // ... takes the chunk again
	CHECK(has(g_ctx[1]->tasks[1], {0, 0, 0}) && has(g_ctx[0]->tasks[1], {-1, 0, 0}));
	for (int r = 0; r < 2; ++r)
		for (const P3 &t : g_ctx[r]->tasks[1])
			CHECK(disp.Owner(t[0]) == r);

	// a stone block in the middle of GPU 1's slab changes a sunlight height: 27 remeshes, none across the cut (chunk (1,0,0):
	// the neighbours' x is 0..2, all GPU 1's), and GPU 0 hears nothing
	sunk.clear(), asked.clear();
	edits = {{{1, 0, 0}, (uint16_t)(16 + (20 * 32 + 16) * 32), 1}};
	world.at({1, 0, 0})->blocks[edits[0].index] = 1;
	disp.SetBlocks(
	    edits, [&](const D::SunlightChange &s) { world.at({s.chunk.x, s.chunk.y, s.chunk.z})->heights[s.index] = s.new_height, ++n_sun; },
	    [&](const ChunkPos3 &p, bool) {
		    asked.push_back(p);
		    if (is_task(p))
			    disp.Submit(gather(p.x, p.y, p.z), true);
	    });
	CHECK(n_sun == 1 && world.at({1, 0, 0})->heights[16 + 16 * 32] == 21 && g_ctx[0]->edits.size() == 1 && g_ctx[1]->edits.size() == 1);
	// GPU 1's ctx sees x = -1 .. 3 whole for task positions x = 0 .. 2: it reports 0..2, nothing is foreign
	CHECK(asked.size() == 27);
	disp.Flush(sink);
	CHECK(g_ctx[0]->tasks.size() == 2 && g_ctx[1]->tasks.size() == 3 && g_ctx[1]->uploads[2].empty() && sunk.size() == 27 && g_violations == 0);

	// a stone block in GPU 1's first plane: the height change remeshes x = -1 tasks too (GPU 0's, derived), and GPU 0 takes the chunk
	sunk.clear(), asked.clear();
	edits = {{{0, 0, 0}, (uint16_t)(16 + (20 * 32 + 16) * 32), 1}};
	world.at({0, 0, 0})->blocks[edits[0].index] = 1;
	disp.SetBlocks(
	    edits, [&](const D::SunlightChange &s) { world.at({s.chunk.x, s.chunk.y, s.chunk.z})->heights[s.index] = s.new_height; },
	    [&](const ChunkPos3 &p, bool) {
		    asked.push_back(p);
		    if (is_task(p))
			    disp.Submit(gather(p.x, p.y, p.z), true);
	    });
	CHECK(asked.size() == 27 && asked_has({-1, 1, -1}) && asked_has({1, -1, 1}));
	disp.Flush(sink);
	CHECK(g_ctx[0]->tasks.size() == 3 && g_ctx[0]->tasks[2].size() == 9);
	if (peer)
		CHECK(g_ctx[0]->uploads[2].empty() && g_ctx[0]->peer_copies.size() == 1 && g_ctx[0]->peer_copies[0].first == g_ctx[1] &&
		      (g_ctx[0]->peer_copies[0].second == std::vector<P3>{P3{0, 0, 0}}));
	else
		CHECK(g_ctx[0]->uploads[2].size() == 1 && g_ctx[0]->uploads[2][0] == (P3{0, 0, 0}));
	CHECK(g_ctx[1]->tasks.size() == 4 && g_ctx[1]->tasks[3].size() == 18 && g_ctx[1]->uploads[3].empty() && sunk.size() == 27 && g_violations == 0);

	// both GPUs' rounds on their own host threads: the sink is called concurrently, every task exactly once
	{
		for (int x = -3; x <= 2; ++x)
			for (int y = -1; y <= 1; ++y)
				for (int z = -1; z <= 1; ++z)
					disp.Submit(gather(x, y, z));
		std::atomic<int> pushed{0}, wrong_rank{0};
		std::mutex mu;
		std::set<P3> seen;
		disp.FlushParallel([&](const ChunkPos3 &p, std::vector<hcm_host::BlockMesh> &&) {
			++pushed;
			std::lock_guard<std::mutex> g(mu);
			wrong_rank += !seen.insert({p.x, p.y, p.z}).second;
		});
		CHECK(pushed == 54 && wrong_rank == 0 && seen.size() == 54 && disp.Pending() == 0 && g_violations == 0);
		CHECK(g_ctx[0]->tasks.back().size() == 27 && g_ctx[1]->tasks.back().size() == 27 && g_ctx[0]->uploads.back().empty() && g_ctx[1]->uploads.back().empty());
	}
	} // peer / host halo

	printf("%s\n", fails ? "FAIL" : "PASS");
	return fails != 0;
}
