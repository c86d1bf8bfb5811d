// GPU test of hcm_host::ShardedMeshDispatcher (hypercraft_b200/host/hcm_host.hpp) on TWO devices through the
// real C ABI: an 8 x 5 x 5 world cut into two x-slabs, one MeshTaskDispatcher (one hcm_ctx) per GPU, both rounds on their own host
// threads (FlushParallel), every pushed std::vector<BlockMesh> compared byte for byte with the CPU oracle (oracle/libhcoracle.so
// via dlopen -- test infrastructure only).  Then edits next to the cut through SetBlocks: the owner applies them on the device,
// the halo holder re-uploads the touched chunks, the remesh fan-out lands on both GPUs, and every remeshed chunk again equals
// the oracle on the host's world.  Prints SKIP and exits 0 when the box has one GPU.
//   usage: test_sharded_gpu <registry.bin> <libhcoracle.so>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <fstream>
#include <map>
#include <set>

#include "../../hypercraft_b200/host/hcm_host.hpp"

struct Pos { int x, y, z; };
struct Chunk { // the accessors of hc::client::Chunk that the mesh task uses (client/include/client/Chunk.hpp:45-66)
	Pos pos;
	std::vector<uint16_t> blocks = std::vector<uint16_t>(32768);
	std::vector<int8_t> heights = std::vector<int8_t>(1024);
	const uint16_t *GetBlockData() const { return blocks.data(); }
	int8_t GetSunlightHeight(uint32_t i) const { return heights[i]; }
	const Pos &GetPosition() const { return pos; }
};

static uint64_t sm64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

int main(int argc, char **argv) {
	if (argc < 3)
		return 2;
	std::ifstream f(argv[1], std::ios::binary);
	std::vector<char> blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	{ // is there a second device?  (no CUDA headers here: ask the library)
		hcm_registry reg{};
		if (hcm_registry_from_blob(blob.data(), blob.size(), &reg))
			return 2;
		hcm_config cfg{};
		cfg.device = 1, cfg.extent[0] = cfg.extent[1] = cfg.extent[2] = 3;
		hcm_ctx *probe = nullptr;
		if (hcm_create(&reg, &cfg, &probe) != HCM_OK) {
			printf("SKIP: no second device (%s)\n", hcm_last_error(nullptr));
			return 0;
		}
		hcm_destroy(probe);
	}
	void *so = dlopen(argv[2], RTLD_NOW);
	if (!so) {
		fprintf(stderr, "dlopen: %s\n", dlerror());
		return 2;
	}
	auto hco_create = (void *(*)(const void *, uint64_t))dlsym(so, "hco_create");
	auto hco_mesh = (int64_t(*)(void *, const uint16_t *const *, const int8_t *const *, uint8_t *, uint64_t))dlsym(so, "hco_mesh");
	auto hco_off = (void (*)(uint32_t, int32_t *))dlsym(so, "hco_neighbour_offset");
	void *oracle = hco_create(blob.data(), blob.size());

	// chunks x in [-4,3], y and z in [-2,2]: rough terrain with water, plants, glowstone and flowing-water cells
	std::map<std::array<int, 3>, std::shared_ptr<Chunk>> world;
	for (int cx = -4; cx <= 3; ++cx)
		for (int cy = -2; cy <= 2; ++cy)
			for (int cz = -2; cz <= 2; ++cz) {
				auto c = std::make_shared<Chunk>();
				c->pos = {cx, cy, cz};
				for (int z = 0; z < 32; ++z)
					for (int x = 0; x < 32; ++x) {
						const int wx = cx * 32 + x, wz = cz * 32 + z;
						const int h = (int)(sm64((uint64_t)((wx >> 3) * 7919 + (wz >> 3))) % 24) - 8 + (int)(sm64((uint64_t)(wx * 31 + wz)) % 3);
						int top = -1;
						for (int y = 0; y < 32; ++y) {
							const int wy = cy * 32 + y;
							const uint64_t r = sm64(((uint64_t)(uint32_t)wx << 40) ^ ((uint64_t)(uint32_t)wy << 20) ^ (uint32_t)wz);
							uint16_t b = 0;
							if (wy < h - 2) b = 1;
							else if (wy < h) b = 3;
							else if (wy == h) b = 4;
							else if (wy <= 0) b = (r % 11 == 0) ? (uint16_t)(15 | ((r >> 8) % 9) << 8) : 14;
							else if (wy == h + 1 && r % 13 == 0) b = (uint16_t)(5 | ((r >> 8) % 4) << 8);
							else if (wy == h + 1 && r % 211 == 0) b = 24;
							c->blocks[x + (y * 32 + z) * 32] = b;
							if (b != 0 && (b & 0xff) != 5) top = y;
						}
						c->heights[x + z * 32] = (int8_t)(top + 1);
					}
				world[{cx, cy, cz}] = c;
			}
	// the reference's sunlight rule across the chunk stack (ChunkFloodSunlightTask.cpp:41-57): a column is lit from above only if the
	// chunk above is lit down to its floor
	for (int cx = -4; cx <= 3; ++cx)
		for (int cz = -2; cz <= 2; ++cz)
			for (int cy = 1; cy >= -2; --cy)
				for (int i = 0; i < 1024; ++i)
					if (world[{cx, cy + 1, cz}]->heights[i] > 0)
						world[{cx, cy, cz}]->heights[i] = 32;

	using D = hcm_host::ShardedMeshDispatcher<Chunk>;
	D disp(blob.data(), blob.size(), {{0, -3, 0}, {1, 0, 3}}, -2, -2, 5, 5); // tasks x in [-3,0) on device 0, [0,3) on device 1
	auto gather = [&](int x, int y, int z) {
		std::array<std::shared_ptr<Chunk>, 27> a;
		for (uint32_t i = 0; i < 27; ++i) {
			int32_t d[3];
			hco_off(i, d);
			a[i] = world.at({x + d[0], y + d[1], z + d[2]});
		}
		return a;
	};
	auto is_task = [](const hcm_host::ChunkPos3 &p) { return p.x >= -3 && p.x <= 2 && p.y >= -1 && p.y <= 1 && p.z >= -1 && p.z <= 1; };
	std::atomic<int> bad{0}, pushed{0};
	std::atomic<size_t> verts{0};
	std::mutex mu;
	std::multiset<std::array<int, 3>> seen;
	auto sink = [&](const hcm_host::ChunkPos3 &pos, std::vector<hcm_host::BlockMesh> &&meshes) { // called from both ranks' threads
		auto a = gather(pos.x, pos.y, pos.z);
		const uint16_t *bp[27];
		const int8_t *hp[27];
		for (int i = 0; i < 27; ++i)
			bp[i] = a[i]->blocks.data(), hp[i] = a[i]->heights.data();
		std::vector<uint8_t> want((size_t)hco_mesh(oracle, bp, hp, nullptr, 0));
		hco_mesh(oracle, bp, hp, want.data(), want.size());
		std::vector<uint8_t> got;
		auto put32 = [&](uint32_t v) { got.insert(got.end(), (uint8_t *)&v, (uint8_t *)&v + 4); };
		put32((uint32_t)meshes.size());
		for (auto &m : meshes) {
			put32(m.transparent);
			for (int i = 0; i < 3; ++i) put32(m.aabb.min[i]);
			for (int i = 0; i < 3; ++i) put32(m.aabb.max[i]);
			put32((uint32_t)m.vertices.size()), put32((uint32_t)m.indices.size());
		}
		for (auto &m : meshes) {
			got.insert(got.end(), (uint8_t *)m.vertices.data(), (uint8_t *)m.vertices.data() + m.vertices.size() * 8);
			got.insert(got.end(), (uint8_t *)m.indices.data(), (uint8_t *)m.indices.data() + m.indices.size() * 2);
			while (got.size() & 3) got.push_back(0);
			verts += m.vertices.size();
		}
		bad += got != want;
		++pushed;
		std::lock_guard<std::mutex> g(mu);
		seen.insert({pos.x, pos.y, pos.z});
	};
	// world load: 54 tasks, 27 per GPU, the planes x = -1 and x = 0 uploaded to both
	for (int x = -3; x <= 2; ++x)
		for (int y = -1; y <= 1; ++y)
			for (int z = -1; z <= 1; ++z)
				disp.Submit(gather(x, y, z));
	disp.FlushParallel(sink);
	int fails = 0, fails_early = 0;
	if (pushed != 54 || std::set<std::array<int, 3>>(seen.begin(), seen.end()).size() != 54 || bad) {
		printf("world load: %d pushed, %d mismatches\n", pushed.load(), bad.load());
		++fails;
	}
	// edits on both sides of the cut: a shaft and a lamp in chunk (-1,0,0) (device 0's last plane), stone and water in (0,0,0) / (0,-1,0)
	std::vector<D::SetBlock> edits;
	for (int y = 0; y < 32; ++y)
		edits.push_back({{-1, 0, 0}, (uint16_t)(30 + (y * 32 + 9) * 32), 0});
	edits.push_back({{-1, 0, 0}, (uint16_t)(31 + (3 * 32 + 9) * 32), 24});
	edits.push_back({{0, 0, 0}, (uint16_t)(0 + (31 * 32 + 5) * 32), 1});
	edits.push_back({{0, -1, 0}, (uint16_t)(1 + (31 * 32 + 9) * 32), 14});
	for (auto &e : edits)
		world.at({e.chunk.x, e.chunk.y, e.chunk.z})->blocks[e.index] = e.block; // Chunk::SetBlock on the host's chunks (the authority)
	int n_sun = 0, n_remesh = 0, n_high = 0;
	seen.clear();
	pushed = 0;
	hcm_stats before[2], after[2];
	for (int r = 0; r < 2; ++r)
		hcm_get_stats(disp.Rank(r).Context(), &before[r]);
	const uint32_t changed = disp.SetBlocks(
	    edits, [&](const D::SunlightChange &s) { world.at({s.chunk.x, s.chunk.y, s.chunk.z})->heights[s.index] = s.new_height, ++n_sun; },
	    [&](const hcm_host::ChunkPos3 &p, bool high) {
		    if (is_task(p))
			    disp.Submit(gather(p.x, p.y, p.z), high), ++n_remesh, n_high += high;
	    });
	disp.FlushParallel(sink);
	for (int r = 0; r < 2; ++r)
		hcm_get_stats(disp.Rank(r).Context(), &after[r]);
	// the touched chunks reached the other GPU device to device (hcm_copy_chunks_from), not through the host: each GPU received whole
	// chunks from its neighbour and far less than one chunk's 64 KB from the host during the round
	const uint64_t p2p[2] = {after[0].p2p_bytes - before[0].p2p_bytes, after[1].p2p_bytes - before[1].p2p_bytes};
	const uint64_t h2d[2] = {after[0].h2d_bytes - before[0].h2d_bytes, after[1].h2d_bytes - before[1].h2d_bytes};
	printf("halo replication: GPU 0 received %llu B from GPU 1 and %llu B from the host, GPU 1 %llu B from GPU 0 and %llu B from the host\n",
	       (unsigned long long)p2p[0], (unsigned long long)h2d[0], (unsigned long long)p2p[1], (unsigned long long)h2d[1]);
	if (p2p[0] < 66560 || p2p[1] < 66560 || h2d[0] > 32768 || h2d[1] > 32768)
		++fails_early;
	printf("edits across the cut: %u blocks changed, %d sunlight heights, %d remesh tasks (%d high priority), %d pushed\n", changed, n_sun, n_remesh, n_high,
	       pushed.load());
	bool both_sides = false, left = false, right = false;
	for (auto &p : seen)
		left |= p[0] < 0, right |= p[0] >= 0;
	both_sides = left && right;
	if (!changed || !n_sun || pushed != n_remesh || !both_sides || !n_high || bad)
		++fails;
	fails += fails_early;
	printf("%s: %zu vertices, %d mismatches\n", fails ? "FAIL" : "PASS", verts.load(), bad.load());
	return fails ? 1 : 0;
}
