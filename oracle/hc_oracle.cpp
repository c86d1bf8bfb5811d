// TEST INFRASTRUCTURE -- not part of the shipped product.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product (libhcm.so) never does.
//
// CPU restatement ("port") of HyperCraft's client mesh task, written from the reference's sources in plain
// C++17 with no reference headers.  It follows the REFERENCE'S algorithm (sequential FIFO flood, sequential
// greedy merge) so that it is an independent check of the CUDA path, which uses a different formulation
// (bit-parallel level-synchronous flood, warp-parallel merge).
//
// PARITY PINNING: the reference holds no tests or golden vectors for this path, so this
// restatement is pinned against the reference itself, compiled unmodified into oracle/_ref/libhcref.so:
// tests/test_oracle_golden.py compares serialised meshes byte for byte on reference-terrain chunks and on
// synthetic stress chunks, and tests/golden/ holds hashes recorded from the reference for the synthetic
// (machine-independent) inputs.
//
// Citations are file:line under /root/reference.
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/hcm.h"

namespace {

constexpr int kN = 32;                   // kChunkSize, common/include/common/Size.hpp:8
constexpr uint32_t kUnitBit = 4;         // BlockVertex::kUnitBitOffset, client/include/client/BlockVertex.hpp:9
constexpr uint32_t kUnit = 1u << kUnitBit;

struct Oracle {
	std::vector<uint8_t> blob;
	hcm_registry reg{};
};

// ---- registry lookups (block/include/block/Block.hpp:41,59-73) --------------------------------------------
struct Blk {
	uint16_t data;
	const hcm_block_prop *p;
	uint8_t id() const { return data & 0xff; }   // Block.hpp:24-33 (x86: m_id is the low byte)
	uint8_t meta() const { return data >> 8; }
	bool indirect_pass() const { return p->transparency != 0; } // GetIndirectLightPass, Block.hpp:70
	bool vertical_pass() const { return p->transparency == 2; } // GetVerticalLightPass, Block.hpp:69
};

// ---- texture helpers (texture/include/texture/BlockTexture.hpp:23-79) -----------------------------------
struct TexOps {
	const hcm_registry *r;
	static uint16_t id(uint16_t t) { return t & 0x1fffu; }
	static uint8_t trans(uint16_t t) { return uint8_t(t >> 13); }
	static bool empty(uint16_t t) { return id(t) == 0; }
	bool transparent(uint16_t t) const { return r->tex_flags[id(t)] & HCM_TEX_TRANSPARENT; }
	bool liquid(uint16_t t) const { return r->tex_flags[id(t)] & HCM_TEX_LIQUID; }
	bool trans_pass(uint16_t t) const { return r->tex_flags[id(t)] & HCM_TEX_TRANS_PASS; }
	// BlockTexture::Show, BlockTexture.hpp:68-76
	bool show(uint16_t a, uint16_t b) const {
		if (empty(a) || a == b)
			return false;
		if (!transparent(a) && !transparent(b))
			return false;
		if (liquid(a) && !empty(b))
			return false;
		return !transparent(a) || transparent(b) || liquid(b);
	}
};

// ---- 27-chunk neighbourhood (client/include/client/Chunk.hpp:57-79, common/include/common/Position.hpp:98-153) ----
struct Hood {
	const hcm_registry *r;
	const uint16_t *const *blocks;
	const int8_t *const *heights;
	static uint32_t nidx(int x, int y, int z) { // GetBlockChunkNeighbourIndex, Position.hpp:142-145
		static const uint32_t L[3] = {1, 2, 0}; // CmpXYZ2NeighbourIndex, Position.hpp:138-141
		int cx = x < 0 ? -1 : (x >= kN ? 1 : 0), cy = y < 0 ? -1 : (y >= kN ? 1 : 0), cz = z < 0 ? -1 : (z >= kN ? 1 : 0);
		return L[cx + 1] * 9u + L[cy + 1] * 3u + L[cz + 1];
	}
	Blk block(int x, int y, int z) const { // GetBlockFromNeighbour, Chunk.hpp:57-59
		uint32_t n = nidx(x, y, z);
		int lx = (x + kN) % kN, ly = (y + kN) % kN, lz = (z + kN) % kN;
		uint16_t d = blocks[n][lx + (ly * kN + lz) * kN]; // InnerIndex3FromPos, Position.hpp:98-100
		return Blk{d, r->props + r->prop_index[d]};
	}
	bool sunlit(int x, int y, int z) const { // GetSunlightFromNeighbour, Chunk.hpp:73-79
		uint32_t n = nidx(x, y, z);
		int lx = (x + kN) % kN, ly = (y + kN) % kN, lz = (z + kN) % kN;
		return ly >= heights[n][lx + lz * kN];
	}
};

// ---- light stage (client/src/ChunkMeshTask.cpp:48-111, client/include/client/BlockLightAlgo.hpp:25-73) ----
constexpr int kE = kN + 30; // 62
inline uint32_t ext_index(int x, int y, int z) { return x + 15 + ((z + 15) + (y + 15) * kE) * kE; } // ChunkMeshTask.cpp:10-12

// IsBorderLightInterfere with Bound {-1..33} (BlockLightAlgo.hpp:32-42, ChunkMeshTask.inl:53-56)
inline bool interfere(int x, int y, int z, uint32_t lvl) {
	const int lo = -1, hi = kN + 1;
	uint32_t dist = 0;
	if (x < lo || x >= hi)
		dist += x < lo ? uint32_t(lo - x) : uint32_t(x - hi);
	if (y < lo || y >= hi)
		dist += y < lo ? uint32_t(lo - y) : uint32_t(y - hi);
	if (z < lo || z >= hi)
		dist += z < lo ? uint32_t(lo - z) : uint32_t(z - hi);
	return lvl > dist;
}
inline bool in_bound_bordered(int x, int y, int z) { // BlockLightAlgo.hpp:25-29 with Border 14
	const int lo = -1 - 14, hi = kN + 1 + 14;
	return lo <= x && x < hi && lo <= y && y < hi && lo <= z && z < hi;
}

struct LightEntry {
	int8_t x, y, z;
	uint8_t lvl;
};

// channel 0 = sunlight (low nibble), 1 = torchlight (high nibble): block/include/block/Light.hpp:20-24
template <int Channel> void propagate(std::queue<LightEntry> &q, const Hood &h, std::vector<uint8_t> &buf) {
	auto get = [&buf](int x, int y, int z) -> uint8_t {
		uint8_t d = buf[ext_index(x, y, z)];
		return Channel ? (d >> 4) : (d & 0xf);
	};
	auto set = [&buf](int x, int y, int z, uint8_t l) {
		uint8_t &d = buf[ext_index(x, y, z)];
		d = Channel ? uint8_t((d & 0x0f) | (l << 4)) : uint8_t((d & 0xf0) | l);
	};
	static const int kStep[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}}; // BlockFace.hpp:13-25
	while (!q.empty()) { // BlockLightAlgo.hpp:48-72
		LightEntry e = q.front();
		q.pop();
		if (e.lvl <= 1)
			continue;
		--e.lvl;
		for (int f = 0; f < 6; ++f) {
			int nx = e.x + kStep[f][0], ny = e.y + kStep[f][1], nz = e.z + kStep[f][2];
			if (!in_bound_bordered(nx, ny, nz))
				continue;
			if (!interfere(nx, ny, nz, e.lvl))
				continue;
			if (e.lvl > get(nx, ny, nz) && h.block(nx, ny, nz).indirect_pass()) {
				set(nx, ny, nz, e.lvl);
				if (e.lvl > 1)
					q.push(LightEntry{(int8_t)nx, (int8_t)ny, (int8_t)nz, e.lvl});
			}
		}
	}
}

void light_stage(const Hood &h, std::vector<uint8_t> &buf) {
	buf.assign(size_t(kE) * kE * kE, 0);
	std::queue<LightEntry> sun_q, torch_q;
	// gather + torch seeds, ChunkMeshTask.cpp:49-59
	for (int y = -15; y < kN + 15; ++y)
		for (int z = -15; z < kN + 15; ++z)
			for (int x = -15; x < kN + 15; ++x) {
				uint8_t torch = h.block(x, y, z).p->light_level;
				uint8_t sun = h.sunlit(x, y, z) ? 15 : 0;
				buf[ext_index(x, y, z)] = uint8_t((torch << 4) | sun);
				if (interfere(x, y, z, torch))
					torch_q.push(LightEntry{(int8_t)x, (int8_t)y, (int8_t)z, torch});
			}
	// sunlight seeds, ChunkMeshTask.cpp:61-87
	const int lo = -15, hi = kN + 14;
	for (int y = -15; y < kN + 15; ++y)
		for (int z = -15; z < kN + 15; ++z)
			for (int x = -15; x < kN + 15; ++x) {
				uint8_t sun = buf[ext_index(x, y, z)] & 0xf;
				if (!interfere(x, y, z, sun))
					continue;
				auto full = [&buf](int xx, int yy, int zz) { return (buf[ext_index(xx, yy, zz)] & 0xf) == 15; };
				if ((x == lo || full(x - 1, y, z)) && (x == hi || full(x + 1, y, z)) && (z == lo || full(x, y, z - 1)) &&
				    (z == hi || full(x, y, z + 1)) && (y == lo || full(x, y - 1, z)) && (y == hi || full(x, y + 1, z)))
					continue;
				sun_q.push(LightEntry{(int8_t)x, (int8_t)y, (int8_t)z, sun});
			}
	propagate<0>(sun_q, h, buf);   // ChunkMeshTask.cpp:89-99
	propagate<1>(torch_q, h, buf); // ChunkMeshTask.cpp:100-111
}

// ---- mesh stage (client/include/client/BlockMeshAlgo.hpp) ----------------------------------------------------
struct Light4 { // BlockMeshAlgo.hpp:18-46
	uint8_t sun[4], torch[4], ao[4];
	bool same(const Light4 &o) const { return !memcmp(sun, o.sun, 4) && !memcmp(torch, o.torch, 4) && !memcmp(ao, o.ao, 4); }
	bool flip() const { // GetFlip, BlockMeshAlgo.hpp:38-43
		return std::abs((int32_t)(ao[0] + 1) * (sun[0] + 1) - (ao[2] + 1) * (sun[2] + 1)) <
		       std::abs((int32_t)(ao[1] + 1) * (sun[1] + 1) - (ao[3] + 1) * (sun[3] + 1));
	}
};

struct Mesh { // client/include/client/BlockVertex.hpp:20-25
	std::vector<uint32_t> vertices; // 2 words per vertex
	std::vector<uint16_t> indices;
	uint32_t aabb_min[3] = {UINT32_MAX, UINT32_MAX, UINT32_MAX}, aabb_max[3] = {0, 0, 0}; // dep/AABB.hpp:11
	bool transparent = false;
	void merge(uint32_t x, uint32_t y, uint32_t z) {
		uint32_t p[3] = {x, y, z};
		for (int i = 0; i < 3; ++i) {
			aabb_min[i] = std::min(aabb_min[i], p[i]);
			aabb_max[i] = std::max(aabb_max[i], p[i]);
		}
	}
	// BlockVertex ctor, BlockVertex.hpp:13-17 (all operands promoted to 32-bit, no masking)
	void push_vertex(uint32_t x10, uint32_t y10, uint32_t z10, uint8_t axis, uint8_t face, uint8_t ao, uint8_t sun, uint8_t torch,
	                 uint16_t tex_id, uint8_t tex_trans) {
		vertices.push_back(x10 | (y10 << 10u) | (z10 << 20u) | ((uint32_t)axis << 30u));
		vertices.push_back(((uint32_t)tex_id - 1u) | ((uint32_t)tex_trans << 10u) | ((uint32_t)face << 13u) | ((uint32_t)ao << 16u) |
		                   ((uint32_t)sun << 18u) | ((uint32_t)torch << 24u));
	}
};

struct Mesher {
	const Oracle &o;
	const Hood &h;
	const std::vector<uint8_t> &light; // 62^3 extended buffer
	TexOps tex;
	std::vector<Mesh> meshes;
	Mesh opaque, transparent;
	std::unordered_map<uint32_t, std::array<uint16_t, 6>> dyn_tex; // m_dynamic_block_textures, BlockMeshAlgo.hpp:487-488

	Mesher(const Oracle &o_, const Hood &h_, const std::vector<uint8_t> &l) : o(o_), h(h_), light(l), tex{&o_.reg} {
		transparent.transparent = true;
	}
	static uint32_t key(int x, int y, int z) { return uint32_t((x + 64) | ((y + 64) << 8) | ((z + 64) << 16)); }
	uint8_t sun_at(int x, int y, int z) const { return light[ext_index(x, y, z)] & 0xf; }
	uint8_t torch_at(int x, int y, int z) const { return light[ext_index(x, y, z)] >> 4; }

	// light4_init, BlockMeshAlgo.hpp:48-163.  The reference's kLookup3v table lists, per face and corner, the two
	// edge-adjacent cells and the diagonal cell of the layer in front of the face; here the same cells are named by
	// the corner's signs along the face's two in-plane axes (lower axis first).
	void light4_init(Light4 *l4, uint8_t face, int x, int y, int z) const {
		static const int8_t kCorner[6][4][2] = {
		    {{-1, -1}, {+1, -1}, {+1, +1}, {-1, +1}}, // +x : (y,z)
		    {{-1, -1}, {-1, +1}, {+1, +1}, {+1, -1}}, // -x : (y,z)
		    {{-1, -1}, {-1, +1}, {+1, +1}, {+1, -1}}, // +y : (x,z)
		    {{-1, -1}, {+1, -1}, {+1, +1}, {-1, +1}}, // -y : (x,z)
		    {{-1, -1}, {+1, -1}, {+1, +1}, {-1, +1}}, // +z : (x,y)
		    {{-1, -1}, {-1, +1}, {+1, +1}, {+1, -1}}, // -z : (x,y)
		};
		const int axis = face >> 1, dir = 1 - ((face & 1) << 1);
		const int pa = axis == 0 ? 1 : 0, qa = axis == 2 ? 1 : 2; // in-plane axes, ascending
		int f[3] = {x, y, z};
		f[axis] += dir; // kLookup1v, BlockMeshAlgo.hpp:74
		Blk front = h.block(f[0], f[1], f[2]);
		const bool d3 = front.vertical_pass() || front.p->collision != 1; // direct_pass[3], :122-125
		for (int v = 0; v < 4; ++v) {
			int c[3][3];
			for (int b = 0; b < 3; ++b)
				memcpy(c[b], f, sizeof(f));
			c[0][qa] += kCorner[face][v][1]; // edge cell along q
			c[1][pa] += kCorner[face][v][0]; // diagonal
			c[1][qa] += kCorner[face][v][1];
			c[2][pa] += kCorner[face][v][0]; // edge cell along p
			bool ip[3], dp[3];
			for (int b = 0; b < 3; ++b) { // :114-120
				Blk blk = h.block(c[b][0], c[b][1], c[b][2]);
				ip[b] = blk.indirect_pass();
				dp[b] = blk.vertical_pass() || blk.p->collision != 1;
			}
			uint32_t ao = ((!dp[0] && !dp[2]) || (!dp[1] && !d3)) ? 0u : (uint32_t)std::max(3 - !dp[0] - !dp[1] - !dp[2] - !d3, 0); // :127-131
			l4->ao[v] = (uint8_t)ao;
			uint32_t counter = 1, ssum = sun_at(f[0], f[1], f[2]), tsum = torch_at(f[0], f[1], f[2]); // :134-136
			if (ip[0] || ip[2]) { // :137-159
				for (int b = 0; b < 3; ++b)
					if (ip[b]) {
						++counter;
						ssum += sun_at(c[b][0], c[b][1], c[b][2]);
						tsum += torch_at(c[b][0], c[b][1], c[b][2]);
					}
			}
			l4->sun[v] = uint8_t((ssum << 2u) / counter); // :160-161
			l4->torch[v] = uint8_t((tsum << 2u) / counter);
		}
	}

	// light4_interpolate, BlockMeshAlgo.hpp:165-196
	static void light4_interpolate(const Light4 &lo, const Light4 &hi, uint8_t du, uint8_t dv, uint8_t dw, uint8_t *ao, uint8_t *sun,
	                               uint8_t *torch) {
		auto lerp = [](uint32_t a, uint32_t b, uint8_t t) -> uint32_t { return a * uint32_t((int32_t)kUnit - (int32_t)t) + b * (uint32_t)t; };
		auto ceil_div = [](uint32_t a, uint32_t b) -> uint32_t { return a / b + (a % b ? 1 : 0); };
		auto tri = [&](const uint8_t *l, const uint8_t *h) {
			return lerp(lerp(lerp(l[0], l[1], du), lerp(l[3], l[2], du), dv), lerp(lerp(h[0], h[1], du), lerp(h[3], h[2], du), dv), dw);
		};
		const uint32_t k3 = kUnit * kUnit * kUnit;
		if (*ao == 4)
			*ao = (uint8_t)ceil_div(tri(lo.ao, hi.ao), k3);
		*sun = (uint8_t)ceil_div(tri(lo.sun, hi.sun), k3);
		*torch = (uint8_t)ceil_div(tri(lo.torch, hi.torch), k3);
	}

	// flow_mesh_gen<Config, Variant>, block/block/public/Liquid.hpp:171-282 (+ get_flow_height :164-169)
	static uint32_t flow_mesh(const hcm_custom_mesh &cm, const Blk nb[8], hcm_mesh_face *faces, uint16_t out_tex[6]) {
		const uint32_t unit = cm.liquid_height_unit;
		auto flow_height = [unit](uint8_t variant) -> uint8_t { return (uint8_t)std::max<uint32_t>(variant * unit, 1u); };
		auto nei_height = [&](const Blk &b) -> int8_t { // :179-184 (raw meta, not the table-clamped variant)
			return b.p->liquid_role == HCM_LIQUID_SOURCE ? (int8_t)16
			                                             : (b.p->liquid_role == HCM_LIQUID_FLOW ? (int8_t)flow_height(b.meta())
			                                                                                    : (b.p->liquid_obstacle ? (int8_t)-1 : (int8_t)0));
		};
		int8_t nhs[8];
		for (int i = 0; i < 8; ++i)
			nhs[i] = nei_height(nb[i]);
		const int8_t nbh = (int8_t)flow_height(cm.liquid_variant);
		auto vh = [&](int a, int b, int c) -> uint8_t { // :198-203
			uint8_t mx = (uint8_t)std::max(std::max(nbh, nhs[a]), std::max(nhs[b], nhs[c]));
			if (mx != 16 && (nbh == 0) + (nhs[a] == 0) + (nhs[b] == 0) + (nhs[c] == 0) >= 3)
				return 0;
			return mx;
		};
		const uint8_t h0 = vh(0, 1, 3), h1 = vh(1, 2, 4), h2 = vh(3, 5, 6), h3 = vh(4, 6, 7);
		const uint16_t ft = cm.liquid_flow_tex;
		if (h0 == 16 && h1 == 16 && h2 == 16 && h3 == 16) { // :210-212
			for (int f = 0; f < 6; ++f)
				out_tex[f] = ft;
			return 0;
		}
		auto face = [&](hcm_mesh_face *f, uint8_t axis, uint8_t bf, const uint8_t v[4][3]) {
			memset(f, 0, sizeof(*f));
			f->axis = axis, f->light_face = bf, f->render_face = bf, f->texture = ft;
			for (int i = 0; i < 4; ++i)
				f->verts[i][0] = v[i][0], f->verts[i][1] = v[i][1], f->verts[i][2] = v[i][2], f->verts[i][3] = 4;
		};
		uint32_t n = 0;
		{
			const uint8_t v[4][3] = {{0, h0, 0}, {0, h2, 16}, {16, h3, 16}, {16, h1, 0}}; // top, :214-226
			face(faces + n++, 1, 2, v);
		}
		const bool left_full = h0 == 16 && h2 == 16, left_empty = h0 == 0 && h2 == 0;
		const bool right_full = h1 == 16 && h3 == 16, right_empty = h1 == 0 && h3 == 0;
		const bool front_full = h2 == 16 && h3 == 16, front_empty = h2 == 0 && h3 == 0;
		const bool back_full = h0 == 16 && h1 == 16, back_empty = h0 == 0 && h1 == 0;
		auto is_liquid = [](const Blk &b) { return b.p->liquid_role != HCM_LIQUID_NONE; };
		if (!left_full && !left_empty && !is_liquid(nb[3])) { // :231-244
			const uint8_t v[4][3] = {{0, h0, 0}, {0, 0, 0}, {0, 0, 16}, {0, h2, 16}};
			face(faces + n++, 0, 1, v);
		}
		if (!right_full && !right_empty && !is_liquid(nb[4])) { // :245-258
			const uint8_t v[4][3] = {{16, 0, 0}, {16, h1, 0}, {16, h3, 16}, {16, 0, 16}};
			face(faces + n++, 0, 0, v);
		}
		if (!back_full && !back_empty && !is_liquid(nb[1])) { // :259-272
			const uint8_t v[4][3] = {{0, 0, 0}, {0, h0, 0}, {16, h1, 0}, {16, 0, 0}};
			face(faces + n++, 2, 5, v);
		}
		if (!front_full && !front_empty && !is_liquid(nb[6])) { // :273-286 (neighbours[6])
			const uint8_t v[4][3] = {{0, h2, 16}, {0, 0, 16}, {16, 0, 16}, {16, h3, 16}};
			face(faces + n++, 2, 4, v);
		}
		out_tex[2] = 0;                     // top,    :287
		out_tex[1] = left_full ? ft : 0;    // left,   :288
		out_tex[0] = right_full ? ft : 0;   // right,  :289
		out_tex[4] = front_full ? ft : 0;   // front,  :290
		out_tex[5] = back_full ? ft : 0;    // back,   :291
		out_tex[3] = ft;                    // bottom, :292
		return n;
	}

	Mesh &stream(bool trans) { return trans ? transparent : opaque; }
	// "if indices would exceed, restart" -- BlockMeshAlgo.hpp:272-279 and :396-403.  cur_vertex is read BEFORE the
	// cut, so the first quad of a continuation mesh is indexed from 65532 (reference behaviour, reproduced).
	uint16_t begin_quad(Mesh &info) {
		uint16_t cur_vertex = (uint16_t)(info.vertices.size() / 2);
		if (cur_vertex + 4 > UINT16_MAX) {
			bool trans = info.transparent;
			meshes.push_back(std::move(info));
			info = Mesh{};
			info.transparent = trans;
		}
		return cur_vertex;
	}

	// generate_custom_mesh, BlockMeshAlgo.hpp:198-306 (swizzle YZX: a1 = y, a2 = z, a3 = x)
	void custom_pass() {
		hcm_mesh_face dyn_faces[HCM_MAX_MESH_FACES];
		for (int y = -1; y <= kN; ++y)
			for (int z = -1; z <= kN; ++z)
				for (int x = -1; x <= kN; ++x) {
					const bool inner = 0 <= x && x < kN && 0 <= y && y < kN && 0 <= z && z < kN;
					Blk b = h.block(x, y, z);
					if (!b.p->custom_mesh)
						continue;
					const hcm_custom_mesh &cm = o.reg.meshes[b.p->custom_mesh - 1];
					uint32_t face_count = cm.face_count;
					const hcm_mesh_face *faces = cm.faces;
					if (cm.dynamic_kind == HCM_DYNAMIC_LIQUID_FLOW) { // :210-231
						static const int kFetch[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}}; // Liquid.hpp:291-301
						Blk nb[8];
						for (int i = 0; i < 8; ++i)
							nb[i] = h.block(x + kFetch[i][0], y, z + kFetch[i][1]);
						std::array<uint16_t, 6> t{};
						face_count = flow_mesh(cm, nb, dyn_faces, t.data());
						dyn_tex[key(x, y, z)] = t; // use_dynamic_texture is always set, Liquid.hpp:209
						faces = dyn_faces;
					}
					if (!inner) // :234-236
						continue;
					const uint32_t base[3] = {(uint32_t)x << kUnitBit, (uint32_t)y << kUnitBit, (uint32_t)z << kUnitBit};
					int cur_light_face = -1;
					uint8_t w_ax = 0, u_ax = 0, v_ax = 0;
					Light4 lo{}, hi{};
					for (uint32_t i = 0; i < face_count; ++i) {
						const hcm_mesh_face &mf = faces[i];
						if (mf.light_face != cur_light_face) { // :247-268
							cur_light_face = mf.light_face;
							w_ax = uint8_t(cur_light_face >> 1);
							u_ax = uint8_t((w_ax + 1) % 3), v_ax = uint8_t((w_ax + 2) % 3);
							if (cur_light_face & 1)
								std::swap(u_ax, v_ax);
							int step = 1 - ((cur_light_face & 1) << 1);
							int op[3] = {x, y, z}, ne[3] = {x, y, z};
							op[w_ax] -= step, ne[w_ax] += step;
							light4_init(&lo, (uint8_t)cur_light_face, op[0], op[1], op[2]);
							if (h.block(ne[0], ne[1], ne[2]).indirect_pass())
								light4_init(&hi, (uint8_t)cur_light_face, x, y, z);
							else
								hi = lo;
						}
						Mesh &info = stream(tex.trans_pass(mf.texture));
						uint16_t cur_vertex = begin_quad(info);
						for (int vi = 0; vi < 4; ++vi) { // :281-293
							const uint8_t *vert = mf.verts[vi];
							info.merge(base[0] + vert[0], base[1] + vert[1], base[2] + vert[2]);
							uint8_t du = vert[u_ax], dv = vert[v_ax], dw = std::min(vert[w_ax], (uint8_t)kUnit);
							if (cur_light_face & 1)
								dw = uint8_t((int32_t)kUnit - (int32_t)dw);
							uint8_t ao = vert[3], sun, torch;
							light4_interpolate(lo, hi, du, dv, dw, &ao, &sun, &torch);
							info.push_vertex(base[0] + vert[0], base[1] + vert[1], base[2] + vert[2], mf.axis, mf.render_face, ao, sun, torch,
							                 TexOps::id(mf.texture), TexOps::trans(mf.texture));
						}
						const uint16_t idx[6] = {0, 1, 2, 0, 2, 3}; // :295-301
						for (uint16_t k : idx)
							info.indices.push_back(uint16_t(cur_vertex + k));
					}
				}
	}

	uint16_t face_tex(int x, int y, int z, uint8_t face) const { // get_tex, BlockMeshAlgo.hpp:321-325
		auto it = dyn_tex.find(key(x, y, z));
		return it == dyn_tex.end() ? h.block(x, y, z).p->tex[face] : it->second[face];
	}

	// generate_regular_mesh_axis<Axis>, BlockMeshAlgo.hpp:308-482.  Swizzle YZX gives (BlockAlgo.hpp:70-81):
	// Axis 0: U = y, V = z;  Axis 1: U = z, V = x;  Axis 2: U = y, V = x;  counter = u*32 + v.
	void axis_pass(int axis) {
		static const int kU[3] = {1, 2, 1}, kV[3] = {2, 0, 0};
		const int ua = kU[axis], va = kV[axis];
		std::vector<uint16_t> tmask[2] = {std::vector<uint16_t>(kN * kN + 1, 0), std::vector<uint16_t>(kN * kN + 1, 0)};
		std::vector<Light4> lmask[2] = {std::vector<Light4>(kN * kN + 1), std::vector<Light4>(kN * kN + 1)};
		const uint8_t fa = uint8_t(axis << 1), fb = uint8_t(fa | 1);
		for (int xa = 0; xa <= kN; ++xa) {
			uint8_t face_mask = 0;
			for (int u = 0; u < kN; ++u)
				for (int v = 0; v < kN; ++v) { // :316-343
					int p[3];
					p[axis] = xa, p[ua] = u, p[va] = v;
					int n[3] = {p[0], p[1], p[2]};
					n[axis] -= 1;
					const uint32_t counter = u * kN + v;
					const uint16_t ta = face_tex(n[0], n[1], n[2], fa), tb = face_tex(p[0], p[1], p[2], fb);
					if (xa != 0 && tex.show(ta, tb)) {
						tmask[0][counter] = ta;
						light4_init(&lmask[0][counter], fa, n[0], n[1], n[2]);
						face_mask |= 1u;
					}
					if (xa != kN && tex.show(tb, ta)) {
						tmask[1][counter] = tb;
						light4_init(&lmask[1][counter], fb, p[0], p[1], p[2]);
						face_mask |= 2u;
					}
				}
			for (int inv = 0; inv < 2; ++inv) { // push_mesh, :345-477
				if (!(face_mask & (1u << inv)))
					continue;
				uint16_t *tm = tmask[inv].data();
				Light4 *lm = lmask[inv].data();
				uint32_t counter = 0;
				for (int u = 0; u < kN; ++u)
					for (int v = 0; v < kN;) {
						const uint16_t qt = tm[counter];
						if (TexOps::empty(qt)) {
							++v, ++counter;
							continue;
						}
						const Light4 ql = lm[counter];
						uint32_t width, height;
						for (width = 1; v + (int)width < kN && qt == tm[counter + width] && ql.same(lm[counter + width]); ++width) // :364-368
							;
						for (height = 1; u + (int)height < kN; ++height) { // :371-379
							bool ok = true;
							for (uint32_t k = 0; k < width && ok; ++k) {
								uint32_t idx = counter + k + height * kN;
								ok = qt == tm[idx] && ql.same(lm[idx]);
							}
							if (!ok)
								break;
						}
						uint32_t p[3];
						p[axis] = xa, p[ua] = u, p[va] = v;
						uint32_t uvec[3] = {0, 0, 0}, vvec[3] = {0, 0, 0};
						if (inv ^ (((ua + 1) % 3 == va) ? 1 : 0)) // :382-390
							uvec[ua] = height, vvec[va] = width;
						else
							uvec[va] = width, vvec[ua] = height;
						Mesh &info = stream(tex.trans_pass(qt)); // :394-395
						uint16_t cur_vertex = begin_quad(info);
						info.merge(p[0] << kUnitBit, p[1] << kUnitBit, p[2] << kUnitBit); // :404-409
						info.merge((p[0] + uvec[0] + vvec[0]) << kUnitBit, (p[1] + uvec[1] + vvec[1]) << kUnitBit,
						           (p[2] + uvec[2] + vvec[2]) << kUnitBit);
						const uint8_t qf = uint8_t((axis << 1) | inv);
						const uint32_t corner[4][3] = {{p[0], p[1], p[2]},
						                               {p[0] + uvec[0], p[1] + uvec[1], p[2] + uvec[2]},
						                               {p[0] + uvec[0] + vvec[0], p[1] + uvec[1] + vvec[1], p[2] + uvec[2] + vvec[2]},
						                               {p[0] + vvec[0], p[1] + vvec[1], p[2] + vvec[2]}};
						for (int c = 0; c < 4; ++c) // :411-432
							info.push_vertex(corner[c][0] << kUnitBit, corner[c][1] << kUnitBit, corner[c][2] << kUnitBit, (uint8_t)axis, qf, ql.ao[c],
							                 ql.sun[c], ql.torch[c], TexOps::id(qt), TexOps::trans(qt));
						static const uint16_t kFlip[6] = {0, 1, 2, 0, 2, 3}, kNoFlip[6] = {1, 2, 3, 0, 1, 3}; // :434-460
						const uint16_t *pat = ql.flip() ? kFlip : kNoFlip;
						for (int k = 0; k < 6; ++k)
							info.indices.push_back(uint16_t(cur_vertex + pat[k]));
						for (uint32_t a = 0; a < height; ++a) // :462-466
							std::fill(tm + counter + kN * a, tm + counter + kN * a + width, (uint16_t)0);
						v += width, counter += width;
					}
			}
		}
	}

	void generate() { // Generate, BlockMeshAlgo.hpp:494-521
		custom_pass();
		axis_pass(0);
		axis_pass(1);
		axis_pass(2);
		if (!opaque.vertices.empty())
			meshes.push_back(std::move(opaque));
		if (!transparent.vertices.empty())
			meshes.push_back(std::move(transparent));
	}
};

size_t serialised_size(const std::vector<Mesh> &ms) {
	size_t s = 4 + ms.size() * 36;
	for (const auto &m : ms)
		s += m.vertices.size() * 4 + ((m.indices.size() * 2 + 3) & ~size_t(3));
	return s;
}
size_t serialise(const std::vector<Mesh> &ms, uint8_t *out) {
	uint8_t *p = out;
	auto put32 = [&p](uint32_t v) { memcpy(p, &v, 4), p += 4; };
	put32((uint32_t)ms.size());
	for (const auto &m : ms) {
		put32(m.transparent ? 1u : 0u);
		for (int i = 0; i < 3; ++i)
			put32(m.aabb_min[i]);
		for (int i = 0; i < 3; ++i)
			put32(m.aabb_max[i]);
		put32((uint32_t)(m.vertices.size() / 2));
		put32((uint32_t)m.indices.size());
	}
	for (const auto &m : ms) {
		memcpy(p, m.vertices.data(), m.vertices.size() * 4), p += m.vertices.size() * 4;
		memcpy(p, m.indices.data(), m.indices.size() * 2), p += m.indices.size() * 2;
		while ((p - out) & 3)
			*p++ = 0;
	}
	return size_t(p - out);
}

} // namespace

extern "C" {

void *hco_create(const void *registry_blob, uint64_t size) {
	auto *o = new Oracle{};
	o->blob.assign((const uint8_t *)registry_blob, (const uint8_t *)registry_blob + size);
	const auto *hd = (const hcm_registry_blob_header *)o->blob.data();
	if (size < sizeof(*hd) || hd->magic != HCM_REGISTRY_MAGIC || hd->version != HCM_REGISTRY_VERSION) {
		delete o;
		return nullptr;
	}
	const uint8_t *p = o->blob.data() + sizeof(*hd);
	o->reg.n_props = hd->n_props, o->reg.n_meshes = hd->n_meshes, o->reg.n_textures = hd->n_textures;
	o->reg.prop_index = p, p += 65536;
	o->reg.props = (const hcm_block_prop *)p, p += sizeof(hcm_block_prop) * hd->n_props;
	o->reg.meshes = (const hcm_custom_mesh *)p, p += sizeof(hcm_custom_mesh) * hd->n_meshes;
	o->reg.tex_flags = p, p += hd->n_textures;
	if (size_t(p - o->blob.data()) > size) {
		delete o;
		return nullptr;
	}
	return o;
}
void hco_destroy(void *o) { delete (Oracle *)o; }

// blocks27 / heights27 in the reference's neighbour order (NeighbourIndex2CmpXYZ, Position.hpp:146-152; 26 = centre).
// Returns the serialised size (layout: see oracle/ref_shim.cpp); writes when cap suffices.
int64_t hco_mesh(void *op, const uint16_t *const *blocks27, const int8_t *const *heights27, uint8_t *out, uint64_t cap) {
	const Oracle &o = *(Oracle *)op;
	Hood h{&o.reg, blocks27, heights27};
	std::vector<uint8_t> light;
	light_stage(h, light);
	Mesher m(o, h, light);
	m.generate();
	size_t need = serialised_size(m.meshes);
	if (out && cap >= need)
		serialise(m.meshes, out);
	return (int64_t)need;
}

// light bytes on [-1,32]^3, index (x+1) + ((z+1) + (y+1)*34)*34
int hco_light34(void *op, const uint16_t *const *blocks27, const int8_t *const *heights27, uint8_t *out) {
	const Oracle &o = *(Oracle *)op;
	Hood h{&o.reg, blocks27, heights27};
	std::vector<uint8_t> light;
	light_stage(h, light);
	for (int y = -1; y <= kN; ++y)
		for (int z = -1; z <= kN; ++z)
			for (int x = -1; x <= kN; ++x)
				out[(x + 1) + ((z + 1) + (y + 1) * 34) * 34] = light[ext_index(x, y, z)];
	return 0;
}

// native liquid generator, same contract as hcref_dynamic_mesh_eval
int hco_dynamic_mesh_eval(void *op, uint16_t block, const uint16_t *neighbours, hcm_mesh_face *out_faces, uint16_t *out_tex,
                          int *out_dynamic_texture) {
	const Oracle &o = *(Oracle *)op;
	const hcm_block_prop *p = o.reg.props + o.reg.prop_index[block];
	if (!p->custom_mesh || o.reg.meshes[p->custom_mesh - 1].dynamic_kind != HCM_DYNAMIC_LIQUID_FLOW)
		return -1;
	Blk nb[8];
	for (int i = 0; i < 8; ++i)
		nb[i] = Blk{neighbours[i], o.reg.props + o.reg.prop_index[neighbours[i]]};
	*out_dynamic_texture = 1;
	return (int)Mesher::flow_mesh(o.reg.meshes[p->custom_mesh - 1], nb, out_faces, out_tex);
}

// NeighbourIndex2CmpXYZ, Position.hpp:146-152
void hco_neighbour_offset(uint32_t idx, int32_t *dxyz) {
	static const int kRev[3] = {1, -1, 0};
	dxyz[2] = kRev[idx % 3u];
	idx /= 3u;
	dxyz[1] = kRev[idx % 3u];
	dxyz[0] = kRev[idx / 3u];
}

// FNV-1a 64 of a byte range -- the digest oracle/ref_shim.cpp (hcref_mesh_batch, per_chunk[.][3]) takes of a chunk's serialised meshes
uint64_t hco_fnv1a(const void *data, uint64_t n) {
	const uint8_t *p = (const uint8_t *)data;
	uint64_t h = 1469598103934665603ull;
	for (uint64_t i = 0; i < n; ++i)
		h = (h ^ p[i]) * 1099511628211ull;
	return h;
}

// Checker for whole batches: the arrays of an hcm_batch_view (include/hcm.h) -> per task {n_meshes, n_vertices, n_indices, FNV-1a of the
// task's meshes in the serialised layout of oracle/ref_shim.cpp}, i.e. the row hcref_mesh_batch writes for the same chunk.  Lets bench.py
// and the tests compare EVERY chunk of a batch with the reference, not a sample.
void hco_digest_batch_view(const hcm_task_result *tasks, uint64_t n_tasks, const hcm_mesh_desc *meshes, const uint64_t *vertices,
                           const uint16_t *indices, uint64_t *out /* n_tasks x 4 */, int n_threads) {
	auto work = [&](uint64_t lo, uint64_t hi) {
		std::vector<uint8_t> buf;
		for (uint64_t t = lo; t < hi; ++t) {
			const hcm_mesh_desc *ms = meshes + tasks[t].first_mesh;
			const uint32_t n = tasks[t].n_meshes;
			size_t size = 4 + (size_t)n * 36;
			uint64_t nv = 0, ni = 0;
			for (uint32_t k = 0; k < n; ++k) {
				size += (size_t)ms[k].vertex_count * 8 + (((size_t)ms[k].index_count * 2 + 3) & ~size_t(3));
				nv += ms[k].vertex_count, ni += ms[k].index_count;
			}
			buf.assign(size, 0);
			uint8_t *p = buf.data();
			auto put32 = [&p](uint32_t v) { memcpy(p, &v, 4), p += 4; };
			put32(n);
			for (uint32_t k = 0; k < n; ++k) {
				put32(ms[k].transparent ? 1u : 0u);
				for (int i = 0; i < 3; ++i)
					put32(ms[k].aabb_min[i]);
				for (int i = 0; i < 3; ++i)
					put32(ms[k].aabb_max[i]);
				put32(ms[k].vertex_count);
				put32(ms[k].index_count);
			}
			for (uint32_t k = 0; k < n; ++k) {
				memcpy(p, vertices + ms[k].vertex_offset, (size_t)ms[k].vertex_count * 8), p += (size_t)ms[k].vertex_count * 8;
				memcpy(p, indices + ms[k].index_offset, (size_t)ms[k].index_count * 2), p += (size_t)ms[k].index_count * 2;
				while ((p - buf.data()) & 3)
					*p++ = 0;
			}
			out[4 * t + 0] = n, out[4 * t + 1] = nv, out[4 * t + 2] = ni, out[4 * t + 3] = hco_fnv1a(buf.data(), size);
		}
	};
	if (n_threads < 1)
		n_threads = 1;
	std::vector<std::thread> ts;
	const uint64_t per = (n_tasks + n_threads - 1) / n_threads;
	for (int i = 1; i < n_threads; ++i)
		if ((uint64_t)i * per < n_tasks)
			ts.emplace_back(work, (uint64_t)i * per, std::min(n_tasks, (uint64_t)(i + 1) * per));
	work(0, std::min(n_tasks, per));
	for (auto &t : ts)
		t.join();
}

} // extern "C"
