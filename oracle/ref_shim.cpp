// TEST INFRASTRUCTURE -- not part of the shipped product.
//
// C shim over the UNMODIFIED reference sources under /root/reference, compiled in place by
// oracle/Makefile into oracle/_ref/libhcref.so.  Nothing here restates the mesher: it only
//   * builds reference `hc::client::Chunk` objects from flat arrays,
//   * calls the reference's own `ChunkTaskRunner<ChunkTaskType::kMesh>::Run`
//     (client/src/ChunkMeshTask.cpp:41-127) and captures `PushChunkMesh` through the shadow
//     header oracle/shadow/client/WorldRenderer.hpp,
//   * calls the reference's own `DefaultTerrain::Generate` (client/src/DefaultTerrain.cpp:10-164)
//     to produce benchmark inputs,
//   * drives the reference's own block-edit chain -- ChunkTaskRunner<kSetBlock / kFloodSunlight / kSetSunlight / kMesh>::Run, popped
//     and run by the reference's own ChunkTaskPool::Run on one thread (client/src/ChunkSetBlockTask.cpp,
//     ChunkFloodSunlightTask.cpp, ChunkSetSunlightTask.cpp, ChunkTaskPool.cpp, ChunkPool.cpp, World.cpp, all unmodified) -- behind a
//     ClientBase that does nothing (client/include/client/ClientBase.hpp:15-28: four pure virtuals), to pin hcm_set_blocks,
//   * flattens the reference's block/texture registry (block/include/block/Block.hpp:41,
//     texture/include/texture/BlockTexture.hpp) into the `hcm_registry` blob of include/hcm.h.
//
// Only tests/, __graft_entry__.smoke() and bench.py (input generation, cpu_baseline and
// --impl reference) may load this library.
#include <client/ChunkTaskPool.hpp>

#include <client/BlockLightAlgo.hpp>
#include <client/BlockMeshAlgo.hpp>
#include <client/ClientBase.hpp>
#include <client/DefaultTerrain.hpp>
#include <client/World.hpp>
#include <client/WorldRenderer.hpp>

#include <block/Block.hpp>
// the concrete liquid trait is needed to recognise the dynamic-mesh function pointers
#include <public/Trait.hpp>
#include <public/Water.hpp>

#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <thread>
#include <unordered_map>

#include "../include/hcm.h"

using namespace hc;
using namespace hc::client;

namespace {

struct PosKey {
	int16_t x, y, z;
	bool operator==(const PosKey &r) const { return x == r.x && y == r.y && z == r.z; }
};
struct PosHash {
	size_t operator()(const PosKey &k) const {
		return (size_t)(uint16_t)k.x * 73856093u ^ (size_t)(uint16_t)k.y * 19349663u ^ (size_t)(uint16_t)k.z * 83492791u;
	}
};

struct RefWorld {
	std::shared_ptr<World> world;
	std::shared_ptr<WorldRenderer> renderer; // only used by single-threaded hcref_mesh
	std::unordered_map<PosKey, std::shared_ptr<Chunk>, PosHash> chunks;
};

struct RefRunner {
	std::shared_ptr<World> world;
	std::shared_ptr<WorldRenderer> renderer;
	std::unique_ptr<ChunkTaskRunner<ChunkTaskType::kMesh>> runner;
	RefRunner() {
		world = World::Create(1, 2);
		renderer = WorldRenderer::Create(world);
		runner = std::make_unique<ChunkTaskRunner<ChunkTaskType::kMesh>>();
	}
	void Run(std::array<std::shared_ptr<Chunk>, 27> &&chunks) {
		runner->Run(const_cast<ChunkTaskPool *>(&world->GetChunkTaskPool()),
		            ChunkTaskRunnerData<ChunkTaskType::kMesh>{std::move(chunks)});
	}
};

static bool gather27(RefWorld *w, const int16_t pos[3], std::array<std::shared_ptr<Chunk>, 27> &out) {
	for (uint32_t i = 0; i < 27; ++i) {
		ChunkPos3 d;
		Chunk::NeighbourIndex2CmpXYZ(i, glm::value_ptr(d));
		auto it = w->chunks.find(PosKey{int16_t(pos[0] + d.x), int16_t(pos[1] + d.y), int16_t(pos[2] + d.z)});
		if (it == w->chunks.end())
			return false;
		out[i] = it->second;
	}
	return true;
}

// Serialised mesh list layout (shared with oracle/hc_oracle.cpp and the tests):
//   u32 n_meshes
//   n_meshes x { u32 transparent, u32 aabb_min[3], u32 aabb_max[3], u32 n_vertices, u32 n_indices }
//   then for each mesh: n_vertices x {u32,u32}, n_indices x u16, zero-padded to 4 bytes
static size_t serialised_size(const std::vector<hc::client::BlockMesh> &meshes) {
	size_t s = 4 + meshes.size() * 36;
	for (const auto &m : meshes) {
		s += m.vertices.size() * 8;
		s += (m.indices.size() * 2 + 3) & ~size_t(3);
	}
	return s;
}
static size_t serialise(const std::vector<hc::client::BlockMesh> &meshes, uint8_t *out) {
	uint8_t *p = out;
	auto put32 = [&p](uint32_t v) {
		memcpy(p, &v, 4);
		p += 4;
	};
	put32((uint32_t)meshes.size());
	for (const auto &m : meshes) {
		put32(m.transparent ? 1u : 0u);
		put32(m.aabb.GetMin().x), put32(m.aabb.GetMin().y), put32(m.aabb.GetMin().z);
		put32(m.aabb.GetMax().x), put32(m.aabb.GetMax().y), put32(m.aabb.GetMax().z);
		put32((uint32_t)m.vertices.size());
		put32((uint32_t)m.indices.size());
	}
	for (const auto &m : meshes) {
		static_assert(sizeof(BlockVertex) == 8);
		memcpy(p, m.vertices.data(), m.vertices.size() * 8);
		p += m.vertices.size() * 8;
		size_t ib = m.indices.size() * 2;
		memcpy(p, m.indices.data(), ib);
		p += ib;
		while ((p - out) & 3)
			*p++ = 0;
	}
	return size_t(p - out);
}

static uint64_t fnv1a(const void *data, size_t n, uint64_t h = 1469598103934665603ull) {
	const uint8_t *p = (const uint8_t *)data;
	for (size_t i = 0; i < n; ++i)
		h = (h ^ p[i]) * 1099511628211ull;
	return h;
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- world -------------------------------------------
void *hcref_world_create() {
	auto *w = new RefWorld{};
	w->world = World::Create(1, 2);
	w->renderer = WorldRenderer::Create(w->world);
	return w;
}
void hcref_world_destroy(void *wp) { delete (RefWorld *)wp; }
uint64_t hcref_world_chunk_count(void *wp) { return ((RefWorld *)wp)->chunks.size(); }

// blocks: 32768 x u16 (index x + (y*32 + z)*32), heights: 1024 x i8 (index x + z*32)
void hcref_world_set_chunk(void *wp, const int16_t pos[3], const uint16_t *blocks, const int8_t *heights) {
	auto *w = (RefWorld *)wp;
	auto chunk = Chunk::Create(ChunkPos3{pos[0], pos[1], pos[2]});
	static_assert(sizeof(block::Block) == 2);
	memcpy((void *)&chunk->GetBlockRef(0u), blocks, 32768 * 2);
	for (uint32_t i = 0; i < 1024; ++i)
		chunk->SetSunlightHeight(i, heights[i]);
	chunk->SetGeneratedFlag();
	w->chunks[PosKey{pos[0], pos[1], pos[2]}] = std::move(chunk);
}
int hcref_world_get_chunk(void *wp, const int16_t pos[3], uint16_t *blocks, int8_t *heights) {
	auto *w = (RefWorld *)wp;
	auto it = w->chunks.find(PosKey{pos[0], pos[1], pos[2]});
	if (it == w->chunks.end())
		return -1;
	memcpy(blocks, it->second->GetBlockData(), 32768 * 2);
	for (uint32_t i = 0; i < 1024; ++i)
		heights[i] = it->second->GetSunlightHeight(i);
	return 0;
}

// bulk variant of hcref_world_get_chunk; returns the number of positions that were missing (left untouched)
int64_t hcref_world_get_chunks(void *wp, const int16_t *pos, uint64_t n, uint16_t *blocks, int8_t *heights) {
	auto *w = (RefWorld *)wp;
	int64_t missing = 0;
	for (uint64_t i = 0; i < n; ++i) {
		auto it = w->chunks.find(PosKey{pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]});
		if (it == w->chunks.end()) {
			++missing;
			continue;
		}
		memcpy(blocks + i * 32768, it->second->GetBlockData(), 32768 * 2);
		for (uint32_t k = 0; k < 1024; ++k)
			heights[i * 1024 + k] = it->second->GetSunlightHeight(k);
	}
	return missing;
}

// Reference terrain: DefaultTerrain(seed).Generate for every listed position, n_threads workers.
// (LocalClient.cpp:37 uses seed 12314524.)  Positions already present are regenerated.
void hcref_world_generate(void *wp, uint32_t seed, const int16_t *pos, uint64_t n, int n_threads) {
	auto *w = (RefWorld *)wp;
	std::vector<std::shared_ptr<Chunk>> fresh(n);
	for (uint64_t i = 0; i < n; ++i)
		fresh[i] = Chunk::Create(ChunkPos3{pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]});
	DefaultTerrain terrain{seed};
	std::atomic<uint64_t> next{0};
	auto work = [&]() {
		for (;;) {
			uint64_t i = next.fetch_add(1);
			if (i >= n)
				break;
			terrain.Generate(fresh[i]);
			fresh[i]->SetGeneratedFlag();
		}
	};
	if (n_threads < 1)
		n_threads = 1;
	std::vector<std::thread> ts;
	for (int t = 1; t < n_threads; ++t)
		ts.emplace_back(work);
	work();
	for (auto &t : ts)
		t.join();
	for (uint64_t i = 0; i < n; ++i)
		w->chunks[PosKey{pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]}] = std::move(fresh[i]);
}

// ---------------------------------------------------------------- mesher ------------------------------------------
// One reference mesh task (unmodified Run). Returns serialised size, or -1 if a neighbour is missing.
// Call hcref_result_copy afterwards to fetch the bytes.
int64_t hcref_mesh(void *wp, const int16_t pos[3]) {
	auto *w = (RefWorld *)wp;
	std::array<std::shared_ptr<Chunk>, 27> chunks;
	if (!gather27(w, pos, chunks))
		return -1;
	thread_local std::unique_ptr<ChunkTaskRunner<ChunkTaskType::kMesh>> runner;
	if (!runner)
		runner = std::make_unique<ChunkTaskRunner<ChunkTaskType::kMesh>>();
	runner->Run(const_cast<ChunkTaskPool *>(&w->world->GetChunkTaskPool()),
	            ChunkTaskRunnerData<ChunkTaskType::kMesh>{std::move(chunks)});
	return (int64_t)serialised_size(w->renderer->captured);
}
int64_t hcref_result_copy(void *wp, uint8_t *out) { return (int64_t)serialise(((RefWorld *)wp)->renderer->captured, out); }

struct hcref_batch_stats {
	double seconds;
	uint64_t chunks, meshes, vertices, indices;
	int32_t threads;
	int32_t missing;
};

// Multithreaded reference mesher: the reference's per-token model (ChunkTaskPool.hpp:187-194) --
// every thread owns one ChunkTaskRunner<kMesh> (+ its own World/renderer sink) and pulls pre-captured
// 27-chunk arrays from an atomic counter.  Scheduling (Pop) is outside the timed region.
// per_chunk (optional, n x 4 u64): {n_meshes, n_vertices, n_indices, fnv1a of the serialised meshes}.
void hcref_mesh_batch(void *wp, const int16_t *pos, uint64_t n, int n_threads, uint64_t *per_chunk,
                      hcref_batch_stats *stats) {
	auto *w = (RefWorld *)wp;
	std::vector<std::array<std::shared_ptr<Chunk>, 27>> tasks(n);
	int missing = 0;
	for (uint64_t i = 0; i < n; ++i)
		if (!gather27(w, pos + i * 3, tasks[i]))
			++missing;
	if (n_threads < 1)
		n_threads = 1;
	std::vector<std::unique_ptr<RefRunner>> runners;
	for (int t = 0; t < n_threads; ++t)
		runners.emplace_back(std::make_unique<RefRunner>());

	std::atomic<uint64_t> next{0}, n_meshes{0}, n_vertices{0}, n_indices{0};
	auto work = [&](int t) {
		RefRunner &r = *runners[t];
		std::vector<uint8_t> buf;
		uint64_t lm = 0, lv = 0, li = 0;
		for (;;) {
			uint64_t i = next.fetch_add(1);
			if (i >= n)
				break;
			if (!tasks[i][26])
				continue;
			r.Run(std::move(tasks[i]));
			const auto &ms = r.renderer->captured;
			uint64_t cv = 0, ci = 0;
			for (const auto &m : ms)
				cv += m.vertices.size(), ci += m.indices.size();
			lm += ms.size(), lv += cv, li += ci;
			if (per_chunk) {
				buf.resize(serialised_size(ms));
				serialise(ms, buf.data());
				per_chunk[i * 4 + 0] = ms.size();
				per_chunk[i * 4 + 1] = cv;
				per_chunk[i * 4 + 2] = ci;
				per_chunk[i * 4 + 3] = fnv1a(buf.data(), buf.size());
			}
		}
		n_meshes += lm, n_vertices += lv, n_indices += li;
	};
	auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> ts;
	for (int t = 1; t < n_threads; ++t)
		ts.emplace_back(work, t);
	work(0);
	for (auto &t : ts)
		t.join();
	auto t1 = std::chrono::steady_clock::now();
	if (stats) {
		stats->seconds = std::chrono::duration<double>(t1 - t0).count();
		stats->chunks = n - missing;
		stats->meshes = n_meshes, stats->vertices = n_vertices, stats->indices = n_indices;
		stats->threads = n_threads;
		stats->missing = missing;
	}
}

// Stage-wise debugging: the light stage of Run (ChunkMeshTask.cpp:48-111) up to the point where
// BlockMeshAlgo starts, returning the [-1,32]^3 light bytes (index (x+1) + ((z+1) + (y+1)*34)*34, value =
// block::Light data = torch<<4 | sun).  The flood itself is the reference's BlockLightAlgo::PropagateLight;
// the two seeding loops are transcribed from ChunkMeshTask.cpp:49-87 because Run() cannot be entered half way.
int hcref_light34(void *wp, const int16_t pos[3], uint8_t *out) {
	auto *w = (RefWorld *)wp;
	std::array<std::shared_ptr<Chunk>, 27> nb;
	if (!gather27(w, pos, nb))
		return -1;
	using LightAlgo = BlockLightAlgo<
	    BlockAlgoConfig<InnerPos1, BlockAlgoBound<InnerPos1>{-1, -1, -1, (InnerPos1)kChunkSize + 1,
	                                                         (InnerPos1)kChunkSize + 1, (InnerPos1)kChunkSize + 1}>,
	    14>;
	constexpr int E = kChunkSize + 30;
	std::vector<block::Light> buf(E * E * E);
	auto idx = [](int x, int y, int z) { return x + 15 + ((z + 15) + (y + 15) * E) * E; };
	LightAlgo::Queue sun_q, torch_q;
	for (InnerPos1 y = -15; y < (InnerPos1)kChunkSize + 15; ++y)
		for (InnerPos1 z = -15; z < (InnerPos1)kChunkSize + 15; ++z)
			for (InnerPos1 x = -15; x < (InnerPos1)kChunkSize + 15; ++x) {
				block::Light light = {};
				uint32_t ni = Chunk::GetBlockNeighbourIndex(x, y, z);
				light.SetTorchlight(nb[ni]->GetBlockFromNeighbour(x, y, z).GetLightLevel());
				light.SetSunlight(nb[ni]->GetSunlightFromNeighbour(x, y, z) ? 15 : 0);
				buf[idx(x, y, z)] = light;
				if (LightAlgo::IsBorderLightInterfere(x, y, z, light.GetTorchlight()))
					torch_q.push({{x, y, z}, light.GetTorchlight()});
			}
	const int lo = -15, hi = kChunkSize + 14;
	for (InnerPos1 y = -15; y < (InnerPos1)kChunkSize + 15; ++y)
		for (InnerPos1 z = -15; z < (InnerPos1)kChunkSize + 15; ++z)
			for (InnerPos1 x = -15; x < (InnerPos1)kChunkSize + 15; ++x) {
				auto s = buf[idx(x, y, z)].GetSunlight();
				if (!LightAlgo::IsBorderLightInterfere(x, y, z, s))
					continue;
				auto full = [&](int xx, int yy, int zz) { return buf[idx(xx, yy, zz)].GetSunlight() == 15; };
				if ((x == lo || full(x - 1, y, z)) && (x == hi || full(x + 1, y, z)) && (z == lo || full(x, y, z - 1)) &&
				    (z == hi || full(x, y, z + 1)) && (y == lo || full(x, y - 1, z)) && (y == hi || full(x, y + 1, z)))
					continue;
				sun_q.push({{x, y, z}, s});
			}
	LightAlgo algo{};
	auto get_block = [&nb](auto x, auto y, auto z) -> block::Block {
		return nb[Chunk::GetBlockNeighbourIndex(x, y, z)]->GetBlockFromNeighbour(x, y, z);
	};
	algo.PropagateLight(
	    &sun_q, get_block, [&](auto x, auto y, auto z) -> block::LightLvl { return buf[idx(x, y, z)].GetSunlight(); },
	    [&](auto x, auto y, auto z, block::LightLvl l) { buf[idx(x, y, z)].SetSunlight(l); });
	algo.PropagateLight(
	    &torch_q, get_block, [&](auto x, auto y, auto z) -> block::LightLvl { return buf[idx(x, y, z)].GetTorchlight(); },
	    [&](auto x, auto y, auto z, block::LightLvl l) { buf[idx(x, y, z)].SetTorchlight(l); });
	for (int y = -1; y <= 32; ++y)
		for (int z = -1; z <= 32; ++z)
			for (int x = -1; x <= 32; ++x)
				out[(x + 1) + ((z + 1) + (y + 1) * 34) * 34] = buf[idx(x, y, z)].GetData();
	return 0;
}

// ---------------------------------------------------------------- registry ----------------------------------------
// Flatten the reference registry into the blob format of include/hcm.h (see hcm_registry_blob_* there).
// Returns the number of bytes required; writes them when cap is large enough.
int64_t hcref_registry_blob(uint8_t *out, uint64_t cap) {
	using namespace hc::block;
	using namespace hc::texture;

	std::vector<hcm_block_prop> props;
	std::vector<hcm_custom_mesh> meshes;
	std::vector<uint8_t> prop_index(65536);
	std::map<const block::BlockMesh *, uint32_t> mesh_ids;
	std::map<std::vector<uint8_t>, uint32_t> prop_ids;

	// Recognise the liquid-flow dynamic generator (Liquid.hpp:171-282).  flow_mesh_gen is `inline static`, so its
	// address differs per translation unit and cannot be compared; the generator is identified by its fetch list
	// (Liquid.hpp:291-301) and its template parameter Variant is recovered by calling it: with eight solid
	// neighbours every corner height equals get_flow_height(Variant) (Liquid.hpp:186-196).
	auto classify_dynamic = [](const block::BlockMesh *m, hcm_custom_mesh *cm) -> int {
		if (!m->p_dynamic_mesh_func)
			return 0;
		static const int8_t kExpect[8][3] = {{-1, 0, -1}, {0, 0, -1}, {1, 0, -1}, {-1, 0, 0},
		                                     {1, 0, 0},   {-1, 0, 1}, {0, 0, 1},  {1, 0, 1}};
		if (m->fetch_neighbour_count != 8)
			return -2;
		for (int i = 0; i < 8; ++i)
			if (m->fetch_neighbours[i].x != kExpect[i][0] || m->fetch_neighbours[i].y != kExpect[i][1] ||
			    m->fetch_neighbours[i].z != kExpect[i][2])
				return -2;
		Block nb[8];
		for (auto &b : nb)
			b = Block{(BlockID)1}; // stone: an obstacle, not a liquid
		BlockMeshFace faces[kBlockMeshMaxFaceCount];
		uint32_t fc = 0;
		texture::BlockTexture tex[6];
		bool dyn = false;
		m->p_dynamic_mesh_func(nb, faces, &fc, tex, &dyn);
		constexpr uint32_t unit = 16 / WaterTrait::kFlowDistance;
		uint32_t h = fc ? faces[0].vertices[0].y : 16;
		cm->dynamic_kind = HCM_DYNAMIC_LIQUID_FLOW;
		cm->liquid_source_id = WaterTrait::kSourceID;
		cm->liquid_flow_id = WaterTrait::kFlowID;
		cm->liquid_flow_tex = BlockTexture{WaterTrait::kFlowTex}.GetData();
		cm->liquid_height_unit = unit;
		cm->liquid_variant = h < unit ? 0 : h / unit;
		cm->liquid_obstacle_rule = HCM_OBSTACLE_ID_NOT_AIR; /* WaterTrait::IsObstacle, Water.hpp:13 */
		return 0;
	};

	for (uint32_t data = 0; data < 65536; ++data) {
		Block b;
		b.SetData((uint16_t)data);
		hcm_block_prop p{};
		for (int f = 0; f < 6; ++f)
			p.tex[f] = b.GetTexture(f).GetData();
		p.transparency = (uint8_t)b.GetTransparency();
		p.light_level = b.GetLightLevel();
		p.collision = (uint8_t)b.GetCollision();
		p.custom_mesh = 0;
		// what flow_mesh_gen<WaterTrait,*>::get_nei_height distinguishes (Liquid.hpp:179-184); one liquid family registered
		p.liquid_role = b.GetID() == WaterTrait::kSourceID
		                    ? HCM_LIQUID_SOURCE
		                    : (b.GetID() == WaterTrait::kFlowID ? HCM_LIQUID_FLOW : HCM_LIQUID_NONE);
		p.liquid_obstacle = WaterTrait::IsObstacle(b) ? 1 : 0;
		if (const block::BlockMesh *m = b.GetCustomMesh()) {
			auto it = mesh_ids.find(m);
			if (it == mesh_ids.end()) {
				hcm_custom_mesh cm{};
				cm.face_count = m->face_count;
				for (uint32_t i = 0; i < m->face_count && i < HCM_MAX_MESH_FACES; ++i) {
					const auto &f = m->faces[i];
					cm.faces[i].axis = f.axis;
					cm.faces[i].light_face = f.light_face;
					cm.faces[i].render_face = f.render_face;
					cm.faces[i].texture = f.texture.GetData();
					for (int v = 0; v < 4; ++v) {
						cm.faces[i].verts[v][0] = f.vertices[v].x;
						cm.faces[i].verts[v][1] = f.vertices[v].y;
						cm.faces[i].verts[v][2] = f.vertices[v].z;
						cm.faces[i].verts[v][3] = f.vertices[v].ao;
					}
				}
				if (int e = classify_dynamic(m, &cm))
					return e; // unknown dynamic generator: cannot be mirrored on the device
				it = mesh_ids.emplace(m, (uint32_t)meshes.size()).first;
				meshes.push_back(cm);
			}
			p.custom_mesh = uint8_t(it->second + 1);
		}
		std::vector<uint8_t> key((const uint8_t *)&p, (const uint8_t *)&p + sizeof(p));
		auto pit = prop_ids.find(key);
		if (pit == prop_ids.end()) {
			if (props.size() >= 256)
				return -4;
			pit = prop_ids.emplace(key, (uint32_t)props.size()).first;
			props.push_back(p);
		}
		prop_index[data] = (uint8_t)pit->second;
	}

	// texture flag table: BlockTexture::IsTransparent/IsLiquid/UseTransparentPass by id
	uint32_t n_tex = BlockTextures::kNone; // placeholder, replaced below
	{
		// number of ids = size of the generated table; probe through the enum's last registered value
		uint32_t max_id = 0;
		for (const auto &p : props)
			for (int f = 0; f < 6; ++f)
				max_id = std::max<uint32_t>(max_id, p.tex[f] & 0x1fffu);
		for (const auto &m : meshes) {
			for (uint32_t i = 0; i < m.face_count; ++i)
				max_id = std::max<uint32_t>(max_id, m.faces[i].texture & 0x1fffu);
			max_id = std::max<uint32_t>(max_id, m.liquid_flow_tex & 0x1fffu);
		}
		n_tex = max_id + 1;
	}
	std::vector<uint8_t> tex_flags(n_tex);
	for (uint32_t id = 0; id < n_tex; ++id) {
		BlockTexture t{(BlockTexID)id};
		tex_flags[id] = (t.IsTransparent() ? HCM_TEX_TRANSPARENT : 0) | (t.IsLiquid() ? HCM_TEX_LIQUID : 0) |
		                (t.UseTransparentPass() ? HCM_TEX_TRANS_PASS : 0);
	}

	hcm_registry_blob_header h{};
	h.magic = HCM_REGISTRY_MAGIC;
	h.version = HCM_REGISTRY_VERSION;
	h.n_props = (uint32_t)props.size();
	h.n_meshes = (uint32_t)meshes.size();
	h.n_textures = n_tex;
	uint64_t need = sizeof(h) + 65536 + props.size() * sizeof(hcm_block_prop) + meshes.size() * sizeof(hcm_custom_mesh) +
	                tex_flags.size();
	if (out && cap >= need) {
		uint8_t *p = out;
		memcpy(p, &h, sizeof(h)), p += sizeof(h);
		memcpy(p, prop_index.data(), 65536), p += 65536;
		memcpy(p, props.data(), props.size() * sizeof(hcm_block_prop)), p += props.size() * sizeof(hcm_block_prop);
		memcpy(p, meshes.data(), meshes.size() * sizeof(hcm_custom_mesh)), p += meshes.size() * sizeof(hcm_custom_mesh);
		memcpy(p, tex_flags.data(), tex_flags.size());
	}
	return (int64_t)need;
}

// Black-box evaluation of a registered dynamic mesh generator (BlockMesh::p_dynamic_mesh_func), used by the tests
// to pin the native restatement: block = Block::GetData() of the voxel, neighbours = fetch_neighbours order.
// out_faces: face_count x hcm_mesh_face; out_tex: 6 x BlockTexture::GetData(). Returns face count, -1 if not dynamic.
int hcref_dynamic_mesh_eval(uint16_t block, const uint16_t *neighbours, hcm_mesh_face *out_faces, uint16_t *out_tex,
                            int *out_dynamic_texture) {
	block::Block b;
	b.SetData(block);
	const block::BlockMesh *m = b.GetCustomMesh();
	if (!m || !m->p_dynamic_mesh_func)
		return -1;
	block::Block nb[block::kBlockDynamicMeshMaxNeighbours];
	for (uint32_t i = 0; i < m->fetch_neighbour_count; ++i)
		nb[i].SetData(neighbours[i]);
	block::BlockMeshFace faces[block::kBlockMeshMaxFaceCount];
	uint32_t fc = m->face_count;
	texture::BlockTexture tex[6];
	bool dyn = false;
	m->p_dynamic_mesh_func(nb, faces, &fc, tex, &dyn);
	for (uint32_t i = 0; i < fc; ++i) {
		out_faces[i].axis = faces[i].axis;
		out_faces[i].light_face = faces[i].light_face;
		out_faces[i].render_face = faces[i].render_face;
		out_faces[i].texture = faces[i].texture.GetData();
		for (int v = 0; v < 4; ++v) {
			out_faces[i].verts[v][0] = faces[i].vertices[v].x;
			out_faces[i].verts[v][1] = faces[i].vertices[v].y;
			out_faces[i].verts[v][2] = faces[i].vertices[v].z;
			out_faces[i].verts[v][3] = faces[i].vertices[v].ao;
		}
	}
	for (int f = 0; f < 6; ++f)
		out_tex[f] = tex[f].GetData();
	*out_dynamic_texture = dyn ? 1 : 0;
	return (int)fc;
}

int hcref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }


// ---------------------------------------------------------------- edit chain ---------------------------------------
// A reference World whose chunk pool holds the caller's chunks, a ClientBase that does nothing, and the reference's own task pool
// run to quiescence on the calling thread.  World declares `friend class LocalClient` (World.hpp:39): a class of that name --
// the real one is not part of this build -- may attach itself as the world's client exactly like LocalClient::Create does.
} // extern "C"
namespace hc::client {
class LocalClient final : public ClientBase {
public:
	static std::shared_ptr<LocalClient> Attach(const std::shared_ptr<World> &world) {
		auto c = std::make_shared<LocalClient>();
		c->m_world_ptr = world;
		world->m_client_weak_ptr = c;
		return c;
	}
	bool IsConnected() override { return true; }
	void LoadChunks(std::span<const ChunkPos3>) override {}                                  // the harness fills the chunks itself
	void SetChunkBlocks(ChunkPos3, std::span<const ChunkSetBlockEntry>) override {}        // server notification: nobody to tell
	void SetChunkSunlights(ChunkPos3, std::span<const ChunkSetSunlightEntry>) override {}
};
} // namespace hc::client
namespace {
struct EditWorld {
	std::shared_ptr<World> world;
	std::shared_ptr<LocalClient> client;
	std::shared_ptr<WorldRenderer> renderer;
	std::vector<ChunkPos3> pushed;
	std::vector<std::vector<hc::client::BlockMesh>> pushed_meshes;
	std::unordered_map<PosKey, uint8_t, PosHash> high; // kMesh requests seen queued with high priority
};
} // namespace
extern "C" {

// load_radius: ChunkPool::Update creates (empty, not yet generated) Chunk objects for the reference's load sphere around the
// origin (client/src/ChunkPool.cpp:9-33); hcref_edit_set_chunk fills the ones the caller wants to exist and marks them generated.
void *hcref_edit_create(int load_radius) {
	auto *w = new EditWorld{};
	w->world = World::Create((ChunkPos1)load_radius, (ChunkPos1)(load_radius + 2));
	w->renderer = WorldRenderer::Create(w->world);
	w->renderer->push_log = &w->pushed, w->renderer->mesh_log = &w->pushed_meshes;
	w->client = LocalClient::Attach(w->world);
	w->world->Start();
	return w;
}
void hcref_edit_destroy(void *wp) { delete (EditWorld *)wp; }
int hcref_edit_set_chunk(void *wp, const int16_t pos[3], const uint16_t *blocks, const int8_t *heights) {
	auto *w = (EditWorld *)wp;
	auto chunk = w->world->GetChunkPool().FindRawChunk(ChunkPos3{pos[0], pos[1], pos[2]});
	if (!chunk)
		return -1; // outside the load sphere
	memcpy((void *)&chunk->GetBlockRef(0u), blocks, 32768 * 2);
	for (uint32_t i = 0; i < 1024; ++i)
		chunk->SetSunlightHeight(i, heights[i]);
	chunk->SetGeneratedFlag();
	return 0;
}
int hcref_edit_get_chunk(void *wp, const int16_t pos[3], uint16_t *blocks, int8_t *heights) {
	auto *w = (EditWorld *)wp;
	auto chunk = w->world->GetChunkPool().FindChunk(ChunkPos3{pos[0], pos[1], pos[2]});
	if (!chunk)
		return -1;
	memcpy(blocks, chunk->GetBlockData(), 32768 * 2);
	for (uint32_t i = 0; i < 1024; ++i)
		heights[i] = chunk->GetSunlightHeight(i);
	return 0;
}

// World::SetBlockBulk's push (World.hpp:124-135: one Push<kSetBlock>(chunk, entries, kLocal) per chunk, in the caller's order) and
// then ChunkTaskPool::Run (ChunkTaskPool.cpp:89-133) on this thread with the worker's producer limits (WorldWorker.cpp:19-22)
// until nothing runs any more.  Tick tasks (kUpdateBlock: water spreading etc.) are never produced: no tick passes.
// Returns the number of mesh tasks that ran (PushChunkMesh calls, in order); out_pos / out_high receive up to `cap` of them
// (high = the request was seen queued with ChunkTaskPriority::kHigh), *n_still_queued the mesh requests that can never run
// because one of their 27 chunks does not exist.
int64_t hcref_edit_apply(void *wp, const hcm_block_edit *edits, uint64_t n, int16_t *out_pos, uint8_t *out_high, uint64_t cap, uint64_t *n_still_queued) {
	auto *w = (EditWorld *)wp;
	auto *pool = const_cast<ChunkTaskPool *>(&w->world->GetChunkTaskPool());
	w->pushed.clear(), w->pushed_meshes.clear(), w->high.clear();
	std::vector<ChunkPos3> order;
	std::unordered_map<PosKey, std::vector<ChunkBlockEntry>, PosHash> by_chunk;
	for (uint64_t i = 0; i < n; ++i) {
		const PosKey k{edits[i].chunk[0], edits[i].chunk[1], edits[i].chunk[2]};
		if (!by_chunk.count(k))
			order.push_back(ChunkPos3{k.x, k.y, k.z});
		block::Block b;
		static_assert(sizeof(block::Block) == 2);
		memcpy(&b, &edits[i].block, 2);
		by_chunk[k].push_back(ChunkBlockEntry{.index = (InnerIndex3)edits[i].index, .block = b});
	}
	for (const ChunkPos3 &c : order)
		pool->Push<ChunkTaskType::kSetBlock>(c, std::span<const ChunkBlockEntry>{by_chunk[PosKey{c.x, c.y, c.z}]}, ChunkUpdateType::kLocal);
	ChunkTaskPoolProducerConfig cfg{};
	cfg.max_high_priority_tasks = 64, cfg.max_tick_tasks = 0, cfg.max_tasks = 256;
	ChunkTaskPoolToken token{pool, cfg};
	auto note_priorities = [&]() {
		ChunkTaskPoolLocked locked{pool};
		for (auto it = locked.GetDataMap().begin(); it != locked.GetDataMap().end(); ++it) {
			const auto &mesh = std::get<static_cast<std::size_t>(ChunkTaskType::kMesh)>(it->second);
			if (mesh.IsQueued() && mesh.GetPriority() == ChunkTaskPriority::kHigh)
				w->high[PosKey{it->first.x, it->first.y, it->first.z}] = 1;
		}
	};
	int idle = 0;
	for (int guard = 0; guard < 1000000 && idle < 6; ++guard) {
		note_priorities();
		const size_t pushes = w->pushed.size();
		pool->Run(&token);
		idle = (pool->GetRunningTaskCountApprox() == 0 && w->pushed.size() == pushes) ? idle + 1 : 0;
	}
	uint64_t still = 0;
	{
		ChunkTaskPoolLocked locked{pool};
		for (auto it = locked.GetDataMap().begin(); it != locked.GetDataMap().end(); ++it)
			still += std::get<static_cast<std::size_t>(ChunkTaskType::kMesh)>(it->second).IsQueued();
	}
	if (n_still_queued)
		*n_still_queued = still;
	for (uint64_t i = 0; i < w->pushed.size() && i < cap; ++i) {
		const ChunkPos3 &p = w->pushed[i];
		out_pos[3 * i] = p.x, out_pos[3 * i + 1] = p.y, out_pos[3 * i + 2] = p.z;
		out_high[i] = w->high.count(PosKey{p.x, p.y, p.z}) ? 1 : 0;
	}
	return (int64_t)w->pushed.size();
}
// serialised meshes of the i-th mesh task hcref_edit_apply ran (size query with out == NULL)
int64_t hcref_edit_pushed_mesh(void *wp, uint64_t i, uint8_t *out) {
	auto *w = (EditWorld *)wp;
	if (i >= w->pushed_meshes.size())
		return -1;
	return out ? (int64_t)serialise(w->pushed_meshes[i], out) : (int64_t)serialised_size(w->pushed_meshes[i]);
}


// Mesh tasks through the task pool: Push<kMesh>(pos, high) for every listed position (what ChunkGenerateTask.cpp:38 does after a
// chunk is generated) and ChunkTaskPool::Run to quiescence.  In oracle/_ref/libhcref.so the tasks are run by the reference's own
// ChunkTaskRunner<kMesh>::Run; in oracle/_ref/libhcbind.so (make binding) by hypercraft_b200/host/reference_binding/ChunkMeshTask.cpp,
// the drop-in replacement of that translation unit -- same harness, same sink.  Results: hcref_edit_pushed_mesh.
int64_t hcref_edit_mesh(void *wp, const int16_t *pos, uint64_t n, int high, int16_t *out_pos, uint64_t cap) {
	auto *w = (EditWorld *)wp;
	auto *pool = const_cast<ChunkTaskPool *>(&w->world->GetChunkTaskPool());
	w->pushed.clear(), w->pushed_meshes.clear(), w->high.clear();
	for (uint64_t i = 0; i < n; ++i)
		pool->Push<ChunkTaskType::kMesh>(ChunkPos3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, high != 0);
	ChunkTaskPoolProducerConfig cfg{};
	cfg.max_high_priority_tasks = 64, cfg.max_tick_tasks = 0, cfg.max_tasks = 256;
	ChunkTaskPoolToken token{pool, cfg};
	int idle = 0;
	for (int guard = 0; guard < 10000000 && idle < 6; ++guard) {
		const size_t pushes = w->pushed.size();
		pool->Run(&token);
		idle = (pool->GetRunningTaskCountApprox() == 0 && w->pushed.size() == pushes) ? idle + 1 : 0;
	}
	for (uint64_t i = 0; i < w->pushed.size() && i < cap; ++i)
		out_pos[3 * i] = w->pushed[i].x, out_pos[3 * i + 1] = w->pushed[i].y, out_pos[3 * i + 2] = w->pushed[i].z;
	return (int64_t)w->pushed.size();
}

#ifdef HCREF_WITH_BINDING
} // extern "C"
#include "../hypercraft_b200/host/reference_binding/hcm_binding.hpp"
extern "C" {
// libhcbind.so only: create the process-wide dispatcher of the replacement mesh lane (hcm_binding::Configure) over the given chunk
// window, with the registry exported from the reference tables above.  Returns 0, or the hcm_status of the failure.
int hcref_binding_configure(int device, const int16_t origin[3], const uint16_t extent[3], char *err, uint64_t err_cap) {
	std::vector<uint8_t> blob((size_t)hcref_registry_blob(nullptr, 0));
	hcref_registry_blob(blob.data(), blob.size());
	try {
		hcm_binding::Configure(blob.data(), blob.size(), device, hcm_host::ChunkPos3{origin[0], origin[1], origin[2]}, {extent[0], extent[1], extent[2]});
	} catch (const hcm_host::Error &e) {
		if (err && err_cap)
			snprintf(err, err_cap, "%s", e.what());
		return e.status;
	}
	return 0;
}
void hcref_binding_shutdown() { hcm_binding::Shutdown(); }
#endif

} // extern "C"
