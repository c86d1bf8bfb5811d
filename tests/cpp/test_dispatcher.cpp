// GPU test of the C++ host mirror (hypercraft_b200/host/hcm_host.hpp): feeds MeshTaskDispatcher the 27-chunk arrays a
// reference worker would get from Pop(), and checks every pushed std::vector<BlockMesh> byte for byte against the CPU
// oracle (oracle/libhcoracle.so, loaded with dlopen -- test infrastructure only).
//   usage: test_dispatcher <registry.bin> <libhcoracle.so>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <fstream>
#include <map>

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
	void *so = dlopen(argv[2], RTLD_NOW);
	if (!so) {
		fprintf(stderr, "dlopen: %s\n", dlerror());
		return 2;
	}
	auto hco_create = (void *(*)(const void *, uint64_t))dlsym(so, "hco_create");
	auto hco_mesh = (int64_t(*)(void *, const uint16_t *const *, const int8_t *const *, uint8_t *, uint64_t))dlsym(so, "hco_mesh");
	auto hco_off = (void (*)(uint32_t, int32_t *))dlsym(so, "hco_neighbour_offset");
	void *oracle = hco_create(blob.data(), blob.size());

	// a 5x5x5 world of rough terrain with water, plants, glowstone and a few flowing-water cells
	std::map<std::array<int, 3>, std::shared_ptr<Chunk>> world;
	for (int cx = -2; cx <= 2; ++cx)
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
							if (b != 0 && b != 5 && (b & 0xff) != 5) top = y;
						}
						c->heights[x + z * 32] = (int8_t)(top + 1);
					}
				world[{cx, cy, cz}] = c;
			}

	hcm_host::MeshTaskDispatcher<Chunk> disp(blob.data(), blob.size(), 0, hcm_host::ChunkPos3{-2, -2, -2}, {5, 5, 5});
	auto gather = [&](int x, int y, int z) {
		std::array<std::shared_ptr<Chunk>, 27> a;
		for (uint32_t i = 0; i < 27; ++i) {
			int32_t d[3];
			hco_off(i, d);
			a[i] = world[{x + d[0], y + d[1], z + d[2]}];
		}
		return a;
	};
	for (int x = -1; x <= 1; ++x)
		for (int y = -1; y <= 1; ++y)
			for (int z = -1; z <= 1; ++z)
				disp.Submit(gather(x, y, z));
	int bad = 0, pushed = 0;
	size_t verts = 0;
	auto sink = [&](const hcm_host::ChunkPos3 &pos, std::vector<hcm_host::BlockMesh> &&meshes) {
		auto a = gather(pos.x, pos.y, pos.z);
		const uint16_t *bp[27];
		const int8_t *hp[27];
		for (int i = 0; i < 27; ++i)
			bp[i] = a[i]->blocks.data(), hp[i] = a[i]->heights.data();
		std::vector<uint8_t> want((size_t)hco_mesh(oracle, bp, hp, nullptr, 0));
		hco_mesh(oracle, bp, hp, want.data(), want.size());
		// serialise what the renderer would receive, in the oracle's layout
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
	};
	disp.Flush(sink);
	// edit: one block changes, the reference would remesh the chunk (+ neighbours passing ChunkShouldRemesh)
	world[{0, 0, 0}] = std::make_shared<Chunk>(*world[{0, 0, 0}]);
	world[{0, 0, 0}]->blocks[16 + (16 * 32 + 16) * 32] = 24;
	disp.MarkDirty({0, 0, 0});
	disp.Submit(gather(0, 0, 0));
	disp.Submit(gather(1, 0, 0));
	disp.Flush(sink);
	// device-side edits: dig a shaft through the terrain of chunk (0,0,0) and drop a lamp and some water next to it; the
	// host chunks are edited in place (same objects: nothing is uploaded again), the sunlight heights come back from the
	// device, and the remesh set it returns is meshed and compared with the oracle on the host's world
	{
		using D = hcm_host::MeshTaskDispatcher<Chunk>;
		std::vector<D::SetBlock> edits;
		for (int y = 0; y < 32; ++y)
			edits.push_back({{0, 0, 0}, (uint16_t)(7 + (y * 32 + 9) * 32), 0});
		edits.push_back({{0, 0, 0}, (uint16_t)(8 + (3 * 32 + 9) * 32), 24});
		edits.push_back({{0, -1, 0}, (uint16_t)(7 + (31 * 32 + 9) * 32), 14});
		edits.push_back({{1, 0, -1}, (uint16_t)(0 + (12 * 32 + 31) * 32), 1});
		for (auto &e : edits)
			world[{e.chunk.x, e.chunk.y, e.chunk.z}]->blocks[e.index] = e.block;
		int n_sun = 0, n_remesh = 0;
		const uint32_t changed = disp.SetBlocks(
		    edits, [&](const D::SunlightChange &s) { world[{s.chunk.x, s.chunk.y, s.chunk.z}]->heights[s.index] = s.new_height, ++n_sun; },
		    [&](const hcm_host::ChunkPos3 &p, bool high) { disp.Submit(gather(p.x, p.y, p.z), high), ++n_remesh; });
		disp.Flush(sink);
		printf("device edits: %u blocks changed, %d sunlight heights changed, %d chunks remeshed\n", changed, n_sun, n_remesh);
		if (!changed || !n_remesh)
			++bad;
	}
	printf("%s: %d tasks pushed, %zu vertices, %d mismatches\n", bad ? "FAIL" : "PASS", pushed, verts, bad);
	return bad ? 1 : 0;
}
