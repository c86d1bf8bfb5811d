// CPU test of the host logic of hcm_host::MeshTaskDispatcher (hypercraft_b200/host/hcm_host.hpp) against a RECORDING FAKE
// of the seven C-ABI calls it makes -- test infrastructure only, linked instead of libhcm.so, it meshes nothing.  Checked:
// which chunks cross the boundary (once per chunk object, again after MarkDirty or when the object changed), that a
// position submitted twice is one task, that high-priority tasks go first, that a task budget leaves the rest pending,
// that every sink call receives the meshes of its own task, that a failed boundary call leaves the round pending, and
// that Submit may be called from many threads while the owner flushes (also clean under -fsanitize=thread).
#include <atomic>
#include <cstdio>
#include <map>
#include <set>
#include <thread>

#include "../../hypercraft_b200/host/hcm_host.hpp"

namespace fake {
struct Call {
	std::vector<std::array<int, 3>> uploaded, tasks;
	std::vector<uint16_t> version; // blocks[1] of every uploaded chunk
};
std::vector<Call> calls;
std::vector<hcm_task_result> tasks;
std::vector<hcm_mesh_desc> meshes;
std::vector<uint64_t> vertices;
std::vector<uint16_t> indices;
bool fail_next = false;
std::vector<std::array<int, 3>> removed;
std::set<std::array<int, 3>> resident; // what the device holds
uint32_t tag(int x, int y, int z) { return (uint32_t)((x + 50) * 10000 + (y + 50) * 100 + (z + 50)); }
} // namespace fake

extern "C" {
struct hcm_ctx {
	int dummy;
};
int hcm_registry_from_blob(const void *, size_t, hcm_registry *out) { return *out = hcm_registry{}, HCM_OK; }
int hcm_create(const hcm_registry *, const hcm_config *, hcm_ctx **out) { return *out = new hcm_ctx{}, HCM_OK; }
void hcm_destroy(hcm_ctx *c) { delete c; }
const char *hcm_last_error(const hcm_ctx *) { return "fake"; }
int hcm_remove_chunks(hcm_ctx *, uint32_t n, const int16_t *pos) {
	for (uint32_t i = 0; i < n; ++i) {
		fake::removed.push_back({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
		fake::resident.erase({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
	}
	return HCM_OK;
}
int hcm_is_resident(const hcm_ctx *, const int16_t pos[3]) { return (int)fake::resident.count({pos[0], pos[1], pos[2]}); }
int hcm_set_blocks(hcm_ctx *, uint32_t, const hcm_block_edit *, hcm_edit_result *out) { return *out = hcm_edit_result{}, HCM_OK; }
// one mesh per task: 4 vertices whose first word is the task position's tag, 6 indices, aabb_min[0] = tag
int hcm_upload_and_mesh(hcm_ctx *, uint32_t n_chunks, const int16_t *pos, const uint16_t *blocks, const int8_t *heights, uint32_t n_tasks,
                        const int16_t *task_pos, hcm_batch_view *out) {
	if (fake::fail_next)
		return fake::fail_next = false, HCM_ERR_CUDA;
	fake::Call c;
	for (uint32_t i = 0; i < n_chunks; ++i) {
		c.uploaded.push_back({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
		c.version.push_back(blocks[(size_t)i * HCM_CHUNK_VOXELS + 1]);
		// the marshalled arrays carry the chunk's own data (block 0 = its tag, height 0 = its x)
		if (blocks[(size_t)i * HCM_CHUNK_VOXELS] != (uint16_t)fake::tag(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]) || heights[(size_t)i * HCM_CHUNK_COLUMNS] != (int8_t)pos[3 * i])
			return HCM_ERR_INVALID;
	}
	for (const auto &u : c.uploaded)
		fake::resident.insert(u);
	fake::tasks.clear(), fake::meshes.clear(), fake::vertices.clear(), fake::indices.clear();
	for (uint32_t t = 0; t < n_tasks; ++t) {
		c.tasks.push_back({task_pos[3 * t], task_pos[3 * t + 1], task_pos[3 * t + 2]});
		for (int dx = -1; dx <= 1; ++dx) // like the device: one missing neighbour fails the WHOLE call
			for (int dy = -1; dy <= 1; ++dy)
				for (int dz = -1; dz <= 1; ++dz)
					if (!fake::resident.count({task_pos[3 * t] + dx, task_pos[3 * t + 1] + dy, task_pos[3 * t + 2] + dz}))
						return HCM_ERR_NOT_RESIDENT;
		const uint32_t g = fake::tag(task_pos[3 * t], task_pos[3 * t + 1], task_pos[3 * t + 2]);
		hcm_mesh_desc d{};
		d.task = t, d.transparent = t & 1, d.vertex_offset = fake::vertices.size(), d.index_offset = fake::indices.size(), d.vertex_count = 4, d.index_count = 6;
		d.aabb_min[0] = g, d.aabb_max[2] = g + 1;
		fake::tasks.push_back({(uint32_t)fake::meshes.size(), 1});
		fake::meshes.push_back(d);
		for (int k = 0; k < 4; ++k)
			fake::vertices.push_back((uint64_t)g | ((uint64_t)k << 32));
		for (uint16_t k : {0, 1, 2, 0, 2, 3})
			fake::indices.push_back(k);
	}
	fake::calls.push_back(c);
	*out = hcm_batch_view{n_tasks, (uint32_t)fake::meshes.size(), fake::vertices.size(), fake::indices.size(), fake::tasks.data(), fake::meshes.data(),
	                      fake::vertices.data(), fake::indices.data()};
	return HCM_OK;
}
}

struct Pos {
	int x, y, z;
};
struct Chunk {
	Pos pos;
	const uint16_t *GetBlockData() const { return blocks.data(); }
	int8_t GetSunlightHeight(uint32_t i) const { return i == 0 ? (int8_t)pos.x : (int8_t)7; }
	const Pos &GetPosition() const { return pos; }
	std::vector<uint16_t> blocks;
	explicit Chunk(Pos p) : pos(p), blocks(HCM_CHUNK_VOXELS, 0) { blocks[0] = (uint16_t)fake::tag(p.x, p.y, p.z); }
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
	std::map<std::array<int, 3>, std::shared_ptr<Chunk>> world;
	for (int x = -2; x <= 2; ++x)
		for (int y = -2; y <= 2; ++y)
			for (int z = -2; z <= 2; ++z)
				world[{x, y, z}] = std::make_shared<Chunk>(Pos{x, y, z});
	auto gather = [&](int x, int y, int z) {
		std::array<std::shared_ptr<Chunk>, 27> a;
		int i = 0; // any fixed order with the centre last (index 26) will do for the host logic
		for (int dx = -1; dx <= 1; ++dx)
			for (int dy = -1; dy <= 1; ++dy)
				for (int dz = -1; dz <= 1; ++dz)
					if (dx || dy || dz)
						a[i++] = world.at({x + dx, y + dy, z + dz});
		a[26] = world.at({x, y, z});
		return a;
	};
	const char blob[4] = {};
	hcm_host::MeshTaskDispatcher<Chunk> disp(blob, sizeof blob, 0, ChunkPos3{-2, -2, -2}, {5, 5, 5});
	std::vector<ChunkPos3> sunk;
	auto sink = [&](const ChunkPos3 &p, std::vector<hcm_host::BlockMesh> &&m) {
		const uint32_t g = fake::tag(p.x, p.y, p.z);
		CHECK(m.size() == 1 && m[0].vertices.size() == 4 && m[0].indices.size() == 6);
		CHECK(m[0].vertices[0].x10_y10_z10_axis2 == g && m[0].vertices[3].tex10_trans3_face3_ao2_sl6_tl6 == 3); // this task's own bytes
		CHECK(m[0].aabb.min[0] == g && m[0].aabb.max[2] == g + 1 && m[0].indices[5] == 3);
		sunk.push_back(p);
	};

	// round 1: two neighbouring tasks share 18 chunks -- every chunk object crosses the boundary once; a repeated request is one task
	disp.Submit(gather(0, 0, 0));
	disp.Submit(gather(1, 0, 0));
	disp.Submit(gather(0, 0, 0));
	CHECK(disp.Pending() == 2);
	disp.Flush(sink);
	CHECK(fake::calls.size() == 1 && fake::calls[0].uploaded.size() == 36 && fake::calls[0].tasks.size() == 2);
	CHECK((std::set<std::array<int, 3>>(fake::calls[0].uploaded.begin(), fake::calls[0].uploaded.end()).size() == 36));
	CHECK(sunk.size() == 2 && sunk[0] == (ChunkPos3{0, 0, 0}) && sunk[1] == (ChunkPos3{1, 0, 0}) && disp.Pending() == 0);
	disp.Flush(sink); // nothing pending: no call
	CHECK(fake::calls.size() == 1);

	// round 2: everything is resident -- nothing is uploaded; an edit (high priority) overtakes two streaming requests
	sunk.clear();
	disp.Submit(gather(1, 0, 0));
	disp.Submit(gather(0, 1, 0));
	disp.Submit(gather(0, 0, 0), true);
	disp.Flush(sink);
	CHECK(fake::calls.size() == 2 && fake::calls[1].tasks.size() == 3);
	CHECK(fake::calls[1].uploaded.size() == 9); // the y = 2 layer of task (0,1,0) was never seen before; the rest is resident
	CHECK(sunk.size() == 3 && sunk[0] == (ChunkPos3{0, 0, 0}) && sunk[1] == (ChunkPos3{1, 0, 0}) && sunk[2] == (ChunkPos3{0, 1, 0}));

	// round 3: MarkDirty and a replaced chunk object both cause exactly that chunk to go up again
	sunk.clear();
	disp.MarkDirty({1, 1, 1});
	world[{-1, 0, 0}] = std::make_shared<Chunk>(*world[{-1, 0, 0}]);
	disp.Submit(gather(0, 0, 0));
	disp.Flush(sink);
	CHECK(fake::calls.size() == 3 && fake::calls[2].uploaded.size() == 2);
	CHECK((std::set<std::array<int, 3>>(fake::calls[2].uploaded.begin(), fake::calls[2].uploaded.end()) == std::set<std::array<int, 3>>{{1, 1, 1}, {-1, 0, 0}}));

	// round 4: a budget of two tasks per flush; the high-priority one and the oldest low one run, the third stays pending
	sunk.clear();
	disp.Submit(gather(0, 0, 0));
	disp.Submit(gather(0, 0, 1));
	disp.Submit(gather(0, -1, 0), true);
	disp.Flush(sink, 2);
	CHECK(sunk.size() == 2 && sunk[0] == (ChunkPos3{0, -1, 0}) && sunk[1] == (ChunkPos3{0, 0, 0}) && disp.Pending() == 1);
	disp.Flush(sink, 2);
	CHECK(sunk.size() == 3 && sunk[2] == (ChunkPos3{0, 0, 1}) && disp.Pending() == 0);
	CHECK(fake::calls.size() == 5 && fake::calls[4].uploaded.empty()); // its chunks went up with the first of the two flushes

	// round 5: the boundary call fails -- Flush throws, nothing is pushed, the round (tasks, priorities, uploads) is pending again
	sunk.clear();
	disp.MarkDirty({0, 0, 0});
	disp.Submit(gather(0, 0, 0));
	disp.Submit(gather(1, 0, 0), true);
	fake::fail_next = true;
	bool threw = false;
	try {
		disp.Flush(sink);
	} catch (const hcm_host::Error &e) {
		threw = e.status == HCM_ERR_CUDA;
	}
	CHECK(threw && sunk.empty() && disp.Pending() == 2 && fake::calls.size() == 5);
	disp.Flush(sink);
	CHECK(fake::calls.size() == 6 && fake::calls[5].uploaded.size() == 1 && fake::calls[5].uploaded[0] == (std::array<int, 3>{0, 0, 0}));
	CHECK(sunk.size() == 2 && sunk[0] == (ChunkPos3{1, 0, 0}) && sunk[1] == (ChunkPos3{0, 0, 0}) && disp.Pending() == 0);

	// round 5b: a chunk is unloaded -- its slot is freed on the device, its own request is dropped, and when it comes back
	// (same object) it crosses the boundary again
	sunk.clear();
	disp.Submit(gather(1, 0, 0));
	disp.Submit(gather(0, 0, 0));
	disp.Unload({1, 0, 0});
	CHECK(fake::removed.size() == 1 && fake::removed[0] == (std::array<int, 3>{1, 0, 0}) && disp.Pending() == 1);
	disp.Submit(gather(0, 0, 0)); // the neighbour is back in the world before the flush
	disp.Flush(sink);
	CHECK(sunk.size() == 1 && sunk[0] == (ChunkPos3{0, 0, 0}));
	CHECK(fake::calls.back().uploaded.size() == 1 && fake::calls.back().uploaded[0] == (std::array<int, 3>{1, 0, 0}));

	// round 5d: the neighbour stays away -- the task that reads it is left out of the round and stays pending (the reference's
	// Pop returns nullopt for that one task, ChunkMeshTask.cpp:29-31), every other task runs; when the chunk is back, so is the task
	sunk.clear();
	disp.Submit(gather(0, 0, 0));
	disp.Submit(gather(-1, 0, 0));
	disp.Submit(gather(-1, -1, 0), true);
	disp.Unload({1, 0, 0}); // read by task (0,0,0) only
	{
		const size_t calls_before = fake::calls.size();
		disp.Flush(sink);
		CHECK(fake::calls.size() == calls_before + 1 && fake::calls.back().tasks.size() == 2);
		CHECK(sunk.size() == 2 && sunk[0] == (ChunkPos3{-1, -1, 0}) && sunk[1] == (ChunkPos3{-1, 0, 0}) && disp.Pending() == 1);
		disp.Flush(sink); // still away: nothing to run, no call, no exception
		CHECK(fake::calls.size() == calls_before + 1 && sunk.size() == 2 && disp.Pending() == 1);
		disp.Submit(gather(1, 0, 0)); // the chunk is in the world again (its own task is requested, and it is uploaded)
		disp.Flush(sink);
		CHECK(sunk.size() == 4 && disp.Pending() == 0);
		CHECK((sunk[2] == (ChunkPos3{0, 0, 0}) && sunk[3] == (ChunkPos3{1, 0, 0})));
	}

	// round 5e: the sink throws at the second task of three -- the first was delivered, the other two are pending again
	sunk.clear();
	disp.Submit(gather(0, 0, 0));
	disp.Submit(gather(0, 0, 1));
	disp.Submit(gather(0, 1, 0));
	{
		int n = 0;
		bool threw2 = false;
		try {
			disp.Flush([&](const ChunkPos3 &p, std::vector<hcm_host::BlockMesh> &&m) {
				if (++n == 2)
					throw std::runtime_error("renderer gone");
				sink(p, std::move(m));
			});
		} catch (const std::runtime_error &) {
			threw2 = true;
		}
		CHECK(threw2 && sunk.size() == 1 && sunk[0] == (ChunkPos3{0, 0, 0}) && disp.Pending() == 2);
		disp.Flush(sink);
		CHECK(sunk.size() == 3 && sunk[1] == (ChunkPos3{0, 0, 1}) && sunk[2] == (ChunkPos3{0, 1, 0}) && disp.Pending() == 0);
	}

	// round 5c: a chunk object is replaced twice while a task that reads it is waiting -- one copy goes up, the latest
	sunk.clear();
	for (uint16_t version : {7, 8}) {
		world[{0, 1, 0}] = std::make_shared<Chunk>(*world.at({0, 1, 0}));
		world.at({0, 1, 0})->blocks[1] = version;
		disp.MarkDirty({0, 1, 0});
		disp.Submit(gather(0, 0, 0));
	}
	disp.Flush(sink);
	CHECK(sunk.size() == 1 && fake::calls.back().uploaded.size() == 1 && fake::calls.back().uploaded[0] == (std::array<int, 3>{0, 1, 0}) &&
	      fake::calls.back().version[0] == 8);

	// round 6: Submit from eight threads (the reference calls Run() from its worker pool) while the owner thread flushes
	{
		sunk.clear();
		const size_t calls_before = fake::calls.size();
		std::set<std::array<int, 3>> uploaded_before;
		for (const auto &c : fake::calls)
			uploaded_before.insert(c.uploaded.begin(), c.uploaded.end());
		for (int x = -2; x <= 2; ++x) // brand-new chunk objects everywhere: each must cross the boundary exactly once more
			for (int y = -2; y <= 2; ++y)
				for (int z = -2; z <= 2; ++z)
					world[{x, y, z}] = std::make_shared<Chunk>(Pos{x, y, z});
		std::atomic<int> running{8};
		std::vector<std::thread> workers;
		for (int w = 0; w < 8; ++w)
			workers.emplace_back([&, w] {
				for (int rep = 0; rep < 20; ++rep)
					for (int i = w; i < 27; i += 4) // every position from two threads, twenty times each
						disp.Submit(gather(i % 3 - 1, i / 3 % 3 - 1, i / 9 - 1), (i + rep) % 5 == 0);
				--running;
			});
		while (running.load())
			disp.Flush(sink, 5);
		for (auto &t : workers)
			t.join();
		while (disp.Pending())
			disp.Flush(sink, 5);
		std::map<std::array<int, 3>, int> up;
		size_t n_tasks = 0;
		for (size_t i = calls_before; i < fake::calls.size(); ++i) {
			for (const auto &c : fake::calls[i].uploaded)
				++up[c];
			n_tasks += fake::calls[i].tasks.size();
			CHECK(fake::calls[i].tasks.size() <= 5);
		}
		bool once = up.size() == 125;
		for (const auto &kv : up)
			once &= kv.second == 1;
		CHECK(once); // 125 new chunk objects, each uploaded once although 27 neighbourhoods x 40 submissions named them
		std::set<std::array<int, 3>> meshed;
		for (const ChunkPos3 &p : sunk)
			meshed.insert({p.x, p.y, p.z});
		CHECK(meshed.size() == 27 && sunk.size() == n_tasks && n_tasks >= 27 && n_tasks < 27 * 40);
	}

	printf("%s: %zu boundary calls\n", fails ? "FAIL" : "PASS", fake::calls.size());
	return fails != 0;
}
