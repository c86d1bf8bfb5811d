// CPU test of hcm_host::MeshRequestQueue (hypercraft_b200/host/hcm_host.hpp): the dedupe / priority rules of the reference's
// per-chunk mesh-task flags (ChunkMeshTask.inl:38-48) and of its producer's pick order (ChunkTaskPool.cpp:16-85).
// Header-only: nothing of libhcm.so is called, so it links without it.
#include <cstdio>

#include "../../hypercraft_b200/host/hcm_host.hpp"

using hcm_host::ChunkPos3;
static int fails = 0;
#define CHECK(c)                                                                                                       \
	do {                                                                                                               \
		if (!(c)) {                                                                                                    \
			printf("FAILED line %d: %s\n", __LINE__, #c);                                                             \
			++fails;                                                                                                   \
		}                                                                                                              \
	} while (0)

int main() {
	hcm_host::MeshRequestQueue q;
	const ChunkPos3 a{0, 0, 0}, b{1, 0, 0}, c{0, -1, 2}, d{5, 5, 5};
	CHECK(q.Push(a) && q.Push(b) && q.Push(c));
	CHECK(!q.Push(a));       // asked for twice before it ran: one task (m_queued is a flag, not a counter)
	CHECK(q.Size() == 3 && q.IsQueued(a) && !q.IsQueued(d));
	CHECK(!q.Push(c, true)); // an edit arrives for a chunk that streaming already queued: sticky high priority
	CHECK(!q.Push(c, false));
	auto r = q.Pop(2);       // high first, then the order of first request
	CHECK(r.size() == 2 && r[0] == c && r[1] == a);
	CHECK(q.Size() == 1 && q.IsQueued(b) && !q.IsQueued(a));
	CHECK(q.Push(a) && q.Push(d, true)); // a is new again after it ran; d overtakes b and a
	CHECK(q.Drop(b) && !q.Drop(b));       // OnUnload
	r = q.Pop();
	CHECK(r.size() == 2 && r[0] == d && r[1] == a && q.Size() == 0);
	CHECK(q.Pop().empty());
	// many positions, every third one high: all come out exactly once, the high ones first, both classes in request order
	for (int i = 0; i < 1000; ++i)
		q.Push(ChunkPos3{(int16_t)(i % 37), (int16_t)(i / 37), 0}, i % 3 == 0), q.Push(ChunkPos3{(int16_t)(i % 37), (int16_t)(i / 37), 0});
	CHECK(q.Size() == 1000);
	size_t got = 0;
	int last_high = -1, last_low = -1;
	bool seen_low = false, ordered = true;
	while (q.Size()) {
		for (const ChunkPos3 &p : q.Pop(64)) {
			const int i = p.y * 37 + p.x;
			if (i % 3 == 0)
				ordered &= !seen_low && i > last_high, last_high = i;
			else
				seen_low = true, ordered &= i > last_low, last_low = i;
			++got;
		}
	}
	CHECK(got == 1000 && ordered);
	printf("%s\n", fails ? "FAIL" : "PASS");
	return fails != 0;
}
