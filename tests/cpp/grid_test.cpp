// C++ caller of the multi-GPU tile-grid driver (include/cpvs_b200.h cpvs_grid_build): what DeferredRenderer::renderWithTiles +
// createShadowTiles + precomputeShadows do (reference src/DeferredRenderer.cpp:150-235), with no NCCL and no Python.
// Builds the same grid on one worker and on several (distinct GPUs when the box has them, else several contexts of GPU 0),
// from a device-generated scene and from a host callback, and requires identical containers and lookups.
//
//   grid_test <tile> <length> <devices...>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cpvs_b200.h"

#define CHECK(call)                                                                          \
	do {                                                                                     \
		int rc_ = (call);                                                                    \
		if (rc_ != CPVS_OK) {                                                                \
			std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cpvs_last_error());     \
			return 1;                                                                        \
		}                                                                                    \
	} while (0)

struct Grid {
	std::vector<uint32_t> dag, cells;
	cpvs_grid_stats stats;
	std::vector<uint8_t> lookups;
};

static int tileSide = 0;
static uint32_t gridLength = 0;

// plane scene of SURVEY.md 8d, written on the host (the fetch path)
static int fetchPlane(void*, uint32_t tx, uint32_t ty, float* out) {
	const float fN = (float)(tileSide * gridLength);
	for (int y = 0; y < tileSide; ++y)
		for (int x = 0; x < tileSide; ++x)
			out[(size_t)y * tileSide + x] = 0.3f + 0.4f * (float)(tx * tileSide + x) / fN + 0.013f * (float)(ty * tileSide + y) / fN;
	return 0;
}

static int buildGrid(const std::vector<int>& devices, const cpvs_grid_desc& desc, const std::vector<float>& points, Grid* out) {
	cpvs_grid* g = nullptr;
	CHECK(cpvs_grid_build(devices.data(), (int)devices.size(), &desc, 1, &g));
	CHECK(cpvs_grid_stats_get(g, &out->stats));
	for (int d = 0; d < (int)devices.size(); ++d) {  // every replica holds the same words
		uint64_t words = 0;
		uint32_t cells = 0;
		CHECK(cpvs_container_info(cpvs_grid_container(g, d), &words, &cells, nullptr, nullptr));
		std::vector<uint32_t> dag(words), grid(cells);
		CHECK(cpvs_container_copy(cpvs_grid_container(g, d), dag.data(), grid.data()));
		if (d == 0) {
			out->dag = dag;
			out->cells = grid;
		} else if (dag != out->dag || grid != out->cells) {
			std::fprintf(stderr, "replica %d differs from replica 0\n", d);
			return 1;
		}
	}
	out->lookups.resize(points.size() / 3);
	CHECK(cpvs_grid_lookup_ndc(g, points.data(), (int64_t)out->lookups.size(), out->lookups.data()));
	CHECK(cpvs_grid_destroy(g));
	return 0;
}

// The worker-level calls a caller with its own scheduling uses: tiles handed out through a callback (the worker asks one tile
// ahead), the cells assembled into a container, the recycled memory given back.
struct TileQueue {
	std::vector<uint32_t> xy;
	size_t next;
};
static int nextTile(void* user, uint32_t* x, uint32_t* y) {
	TileQueue* q = static_cast<TileQueue*>(user);
	if (q->next * 2 >= q->xy.size()) return 0;
	*x = q->xy[2 * q->next];
	*y = q->xy[2 * q->next + 1];
	++q->next;
	return 1;
}
static int buildThroughWorker(int device, const cpvs_grid_desc& desc, std::vector<uint32_t>* dagOut, std::vector<uint32_t>* cellsOut) {
	cpvs_ctx* ctx = nullptr;
	CHECK(cpvs_ctx_create(device, &ctx));
	cpvs_grid_worker* w = nullptr;
	CHECK(cpvs_grid_worker_create(ctx, &desc, &w));
	TileQueue q;
	q.next = 0;
	for (uint32_t y = 0; y < desc.length; ++y)
		for (uint32_t x = 0; x < desc.length; ++x) {
			q.xy.push_back(x);
			q.xy.push_back(y);
		}
	CHECK(cpvs_grid_worker_build_from(w, nextTile, &q));
	const size_t numCells = (size_t)desc.length * desc.length * desc.length;
	std::vector<cpvs_grid_cell> cells(numCells);
	if (cpvs_grid_worker_cells(w, cells.data(), (int)cells.size()) != (int)numCells) {
		std::fprintf(stderr, "worker: %zu cells expected\n", numCells);
		return 1;
	}
	std::vector<cpvs_cell_part> parts(numCells);
	for (const cpvs_grid_cell& c : cells) parts[c.index] = cpvs_cell_part{c.words, c.root_mask, c.device, c.words_device};
	cpvs_container* cont = nullptr;
	CHECK(cpvs_container_assemble(ctx, desc.length, cells[0].num_levels, desc.leafmasks, parts.data(), &cont));
	uint64_t words = 0;
	uint32_t ncells = 0;
	CHECK(cpvs_container_info(cont, &words, &ncells, nullptr, nullptr));
	dagOut->resize(words);
	cellsOut->resize(ncells);
	CHECK(cpvs_container_copy(cont, dagOut->data(), cellsOut->data()));
	CHECK(cpvs_container_destroy(cont));
	CHECK(cpvs_grid_worker_destroy(w));
	CHECK(cpvs_ctx_trim(ctx));
	CHECK(cpvs_ctx_destroy(ctx));
	return 0;
}

int main(int argc, char** argv) {
	if (argc < 4) {
		std::fprintf(stderr, "usage: grid_test <tile> <length> <devices...>\n");
		return 2;
	}
	tileSide = std::atoi(argv[1]);
	gridLength = (uint32_t)std::atoi(argv[2]);
	std::vector<int> devices;
	for (int i = 3; i < argc; ++i) devices.push_back(std::atoi(argv[i]));
	std::vector<float> points(3 * 200000);
	uint32_t s = 777;
	for (float& p : points) {
		s ^= s << 13;
		s ^= s >> 17;
		s ^= s << 5;
		p = (s % 20001) / 10000.f - 1.f;
	}
	const int scenes[] = {CPVS_SCENE_TERRAIN_DEV, CPVS_SCENE_CITY, -1};
	for (int scene : scenes) {
		cpvs_grid_desc desc;
		std::memset(&desc, 0, sizeof(desc));
		desc.length = gridLength;
		desc.tile = tileSide;
		desc.leafmasks = 1;
		desc.scene = scene;
		desc.fetch = scene < 0 ? fetchPlane : nullptr;
		Grid one, many;
		if (buildGrid(std::vector<int>(1, devices[0]), desc, points, &one)) return 1;
		if (buildGrid(devices, desc, points, &many)) return 1;
		if (one.dag != many.dag || one.cells != many.cells || one.lookups != many.lookups) {
			std::fprintf(stderr, "scene %d: %zu workers disagree with one worker (%zu vs %zu words)\n", scene, devices.size(), many.dag.size(), one.dag.size());
			return 1;
		}
		std::vector<uint32_t> workerDag, workerCells;
		if (buildThroughWorker(devices[0], desc, &workerDag, &workerCells)) return 1;
		if (workerDag != one.dag || workerCells != one.cells) {
			std::fprintf(stderr, "scene %d: a worker fed through the callback disagrees with cpvs_grid_build\n", scene);
			return 1;
		}
		unsigned lit = 0;
		for (uint8_t v : one.lookups) lit += v;
		std::printf("scene %2d: %u cells (%u of one word), %llu words, %u / %zu lookups lit; %zu workers: tiles", scene, one.stats.cells,
				one.stats.one_word_cells, (unsigned long long)one.stats.dag_words, lit, one.lookups.size(), devices.size());
		for (size_t d = 0; d < devices.size(); ++d) std::printf(" %u", many.stats.tiles[d]);
		std::printf(", %u moved, build %.2f ms (1 worker %.2f ms)\n", many.stats.moved_tiles, many.stats.build_ms_max, one.stats.build_ms_max);
	}
	std::printf("grid_test ok\n");
	return 0;
}
