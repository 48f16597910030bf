/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's linear-time MSER (row a12 of SURVEY.md section 8).
 * Nothing under compv_b200/ links or calls this file; tests/, __graft_entry__.smoke() and bench.py's CPU legs are its only users.
 *
 * Follows /root/reference/core/ccl/compv_core_ccl_lmser.cxx: the flood of process() :148-368 (edge macro LMSER_CHECK_EDGE :31-50, boundary heap :71-121,
 * ProcessStack :320-366), the collection :370-405; and core/include/compv/core/ccl/compv_core_ccl_lmser_result.h: merge :86-91,
 * collectStableRegions :94-119, computeFinalPoints :122-156, computeVariation :252-262, computeStability :287-307, checkCrit :192-207;
 * bounding boxes core/ccl/compv_core_ccl_lmser_result.cxx:50-88.
 * Pinned against the compiled reference (oracle/_ref) by tests/test_lmser.py: same regions, same order, same point order. */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define HIGHEST 256

typedef struct {
	int sister, child, parent; /* node indices, -1 = none */
	double variation;
	int stable;
	int greyLevel;
	int area;
	int mergesHead, pointsHead; /* pool indices, -1 = empty */
} node_t;

typedef struct {
	node_t* nodes; int nNodes;
	int* ptData; int* ptLink; int nPts;      /* pixel lists (push_front) */
	int* mgData; int* mgLink; int nMg;       /* merge lists (push_front) */
	double oneMinusDiv, oneMinusDivScale;
	int* stableOut; size_t nStable;
	int stride; float strideScale;
	int16_t* points; size_t pointCap; size_t np;
} ctx_t;

static int new_node(ctx_t* c, int level)
{
	node_t* n = &c->nodes[c->nNodes];
	memset(n, 0, sizeof(*n)); /* CompVMemZero (lmser_result.h:226-238) */
	n->sister = n->child = n->parent = -1; n->mergesHead = n->pointsHead = -1;
	n->greyLevel = level;
	return c->nNodes++;
}

static void merge(ctx_t* c, int a, int b) /* a->merge(b) :86-91 */
{
	node_t* A = &c->nodes[a]; node_t* B = &c->nodes[b];
	A->area += B->area;
	B->sister = A->child; A->child = b; B->parent = a;
	c->mgData[c->nMg] = b; c->mgLink[c->nMg] = A->mergesHead; A->mergesHead = c->nMg++;
}

static int check_crit(const ctx_t* c, int n, int area_, double variation_) /* :192-207 */
{
	const node_t* N = &c->nodes[n];
	if (N->area <= area_) return 1;
	if (N->stable && (N->variation < variation_)) return 0;
	for (int ch = N->child; ch >= 0; ch = c->nodes[ch].sister) if (!check_crit(c, ch, area_, variation_)) return 0;
	return 1;
}

static void collect(ctx_t* c, int n) /* :94-119 */
{
	node_t* N = &c->nodes[n];
	if (N->stable) {
		const int min_parent_area = (int)((N->area * c->oneMinusDivScale) + 0.5);
		for (int p = N->parent; p >= 0 && (c->nodes[p].area < min_parent_area) && (N->stable = (!c->nodes[p].stable || (c->nodes[p].variation > N->variation))); p = c->nodes[p].parent) { }
		N->stable = N->stable && check_crit(c, n, (int)((N->area * c->oneMinusDiv) + 0.5), N->variation);
		if (N->stable) c->stableOut[c->nStable++] = n;
	}
	for (int ch = N->child; ch >= 0; ch = c->nodes[ch].sister) collect(c, ch);
}

static void final_points(ctx_t* c, int n) /* :142-155 (the recursive variant the reference compiles) */
{
	const node_t* N = &c->nodes[n];
	for (int k = N->pointsHead; k >= 0; k = c->ptLink[k]) {
		const int16_t y = (int16_t)(c->ptData[k] * c->strideScale);
		const int16_t x = (int16_t)(c->ptData[k] - (y * c->stride));
		if (c->points && c->np < c->pointCap) { c->points[2 * c->np] = x; c->points[2 * c->np + 1] = y; }
		++c->np;
	}
	for (int k = N->mergesHead; k >= 0; k = c->mgLink[k]) final_points(c, c->mgData[k]);
}

/* regions in the reference's order: regionSizes[i] points, regionBoxes[4i..] = {left, top, right, bottom} (inclusive min/max), points = (x, y) int16 pairs concatenated */
ORC_API int orc_ccl_lmser(const uint8_t* img, size_t width_, size_t height_, size_t stride_, int delta, double minArea, double maxArea, double maxVariation, double minDiversity,
	int connectivity, int32_t* regionSizes, int16_t* regionBoxes, size_t regionCap, int16_t* points, size_t pointCap, size_t* regionCount, size_t* pointCount)
{
	if (!img || !width_ || !height_ || stride_ < width_) return 20006;
	if (delta <= 0 || delta > 255 || minArea < 0.0 || minArea > 1.0 || minArea > maxArea || maxArea > 1.0 || maxVariation < 0.0 || maxVariation > 1.0
		|| minDiversity < 0.0 || minDiversity > 1.0 || (connectivity != 4 && connectivity != 8)) return 20006; /* compv_ccl.cxx:76-83 */
	const int width = (int)width_, height = (int)height_, stride = (int)stride_;
	const int b8 = (connectivity == 8), maxEdges = b8 ? 8 : 4;
	int off[8];
	if (b8) { off[0] = 1; off[1] = 1 - stride; off[2] = -stride; off[3] = -stride - 1; off[4] = -1; off[5] = stride - 1; off[6] = stride + 1; off[7] = stride; } /* :177-186 */
	else { off[0] = 1; off[1] = -stride; off[2] = -1; off[3] = stride; }
	const size_t npx = (size_t)width * height;
	ctx_t c; memset(&c, 0, sizeof(c));
	c.nodes = (node_t*)malloc((npx + 2) * sizeof(node_t));
	c.ptData = (int*)malloc(npx * sizeof(int)); c.ptLink = (int*)malloc(npx * sizeof(int));
	c.mgData = (int*)malloc((npx + 2) * sizeof(int)); c.mgLink = (int*)malloc((npx + 2) * sizeof(int));
	uint32_t* bdData = (uint32_t*)malloc((npx + 1) * sizeof(uint32_t)); int* bdLink = (int*)malloc((npx + 1) * sizeof(int));
	const size_t accW = (size_t)stride + 1; /* :214 */
	uint8_t* accAll = (uint8_t*)malloc(((size_t)height + 2) * accW + (size_t)stride * 2 + 16);
	int* stackC = (int*)malloc((HIGHEST + 2) * sizeof(int));
	c.stableOut = (int*)malloc((npx + 2) * sizeof(int));
	if (!c.nodes || !c.ptData || !c.ptLink || !c.mgData || !c.mgLink || !bdData || !bdLink || !accAll || !stackC || !c.stableOut) return 20013;
	memset(accAll, 1, ((size_t)height + 2) * accW + (size_t)stride * 2 + 16);
	uint8_t* acc = accAll + accW; /* :217 */
	for (int y = 0; y < height; ++y) memset(acc + (size_t)y * stride, 0, (size_t)width); /* :218-232 */
	int bdTail[HIGHEST]; for (int i = 0; i < HIGHEST; ++i) bdTail[i] = -1;
	int nBd = 0, sp = 0;
	int current_priority = HIGHEST;
#define BD_PUSH(level, pix) do { bdData[nBd] = (uint32_t)(pix); bdLink[nBd] = bdTail[level]; bdTail[level] = nBd++; } while (0)
	stackC[sp++] = new_node(&c, HIGHEST); /* step 1 :270-271 */
	int current_pixel = 0, current_edge = 0; /* step 2 :275-279 */
	acc[0] = 1;
	int current_level = img[0];
	for (;;) {
		stackC[sp++] = new_node(&c, current_level); /* step 3 :283-284 */
		for (;;) {
			int descended = 0;
			while (current_edge < maxEdges) { /* step 4 :293-305 with LMSER_CHECK_EDGE :31-50 */
				const int neighbor_pixel = current_pixel + off[current_edge];
				if (!acc[neighbor_pixel]) {
					acc[neighbor_pixel] = 1;
					const int neighbor_level = img[neighbor_pixel];
					if (neighbor_level >= current_level) {
						BD_PUSH(neighbor_level, neighbor_pixel);
						if (neighbor_level < current_priority) current_priority = neighbor_level;
					}
					else {
						BD_PUSH(current_level, (uint32_t)current_pixel | ((uint32_t)(current_edge + 1) << 28));
						if (current_level < current_priority) current_priority = current_level;
						current_edge = 0; current_pixel = neighbor_pixel; current_level = neighbor_level;
						descended = 1;
						break;
					}
				}
				++current_edge;
			}
			if (descended) break; /* goto step 3 */
			{ /* step 5 :308-311 */
				node_t* top = &c.nodes[stackC[sp - 1]];
				++top->area;
				c.ptData[c.nPts] = current_pixel; c.ptLink[c.nPts] = top->pointsHead; top->pointsHead = c.nPts++;
			}
			if (current_priority == HIGHEST) goto done; /* step 6 :315-317 */
			{
				const int tail = bdTail[current_priority];
				current_pixel = (int)(bdData[tail] & 0xfffffff);
				current_edge = (int)(bdData[tail] >> 28);
				bdTail[current_priority] = bdLink[tail]; /* pop_back :95-118 */
				if (bdTail[current_priority] < 0) {
					for (; current_priority < HIGHEST && bdTail[current_priority] < 0; ++current_priority) { }
				}
			}
			const int new_level = img[current_pixel]; /* step 7 :327-366 */
			if (new_level != current_level) {
				current_level = new_level;
				do {
					const int top = stackC[--sp];
					if (new_level < c.nodes[stackC[sp - 1]].greyLevel) {
						const int nn = new_node(&c, new_level);
						merge(&c, nn, top);
						stackC[sp++] = nn;
						break;
					}
					merge(&c, stackC[sp - 1], top);
				} while (new_level > c.nodes[stackC[sp - 1]].greyLevel);
			}
		}
	}
done:;
	const int master = stackC[sp - 1]; /* :372 */
	const int input_area = width * height;
	const int min_area_ = (int)(input_area * minArea), max_area_ = (int)(input_area * maxArea);
	c.oneMinusDiv = 1.0 - minDiversity; c.oneMinusDivScale = 1.0 / c.oneMinusDiv;
	for (int i = 0; i < c.nNodes; ++i) { /* computeVariation :252-262 */
		node_t* N = &c.nodes[i];
		const int deltaPlus = N->greyLevel + delta;
		int p = i;
		while (c.nodes[p].parent >= 0 && c.nodes[c.nodes[p].parent].greyLevel <= deltaPlus) p = c.nodes[p].parent;
		N->variation = (c.nodes[p].area - N->area) / (double)N->area;
	}
	for (int i = 0; i < c.nNodes; ++i) { /* computeStability :287-307 */
		node_t* N = &c.nodes[i];
		const int stable_ = (N->parent < 0 || (c.nodes[N->parent].variation >= N->variation)) && (N->variation <= maxVariation) && (min_area_ <= N->area && N->area <= max_area_);
		if (N->child >= 0) {
			if (stable_) for (int ch = N->child; ch >= 0; ch = c.nodes[ch].sister) if (N->variation < c.nodes[ch].variation) { N->stable = 1; break; }
		}
		else N->stable = stable_;
	}
	collect(&c, master); /* :382 */
	c.stride = stride; c.strideScale = 1.f / (float)stride; /* :387 */
	c.points = points; c.pointCap = pointCap; c.np = 0;
	for (size_t i = 0; i < c.nStable; ++i) {
		const size_t first = c.np;
		final_points(&c, c.stableOut[i]);
		if (i < regionCap) {
			if (regionSizes) regionSizes[i] = (int32_t)(c.np - first);
			if (regionBoxes && points && c.np <= pointCap && c.np > first) { /* lmser_result.cxx:60-74 */
				int16_t l = points[2 * first], r = l, t = points[2 * first + 1], b = t;
				for (size_t k = first + 1; k < c.np; ++k) {
					const int16_t x = points[2 * k], y = points[2 * k + 1];
					if (x < l) l = x;
					if (x > r) r = x;
					if (y < t) t = y;
					if (y > b) b = y;
				}
				regionBoxes[4 * i] = l; regionBoxes[4 * i + 1] = t; regionBoxes[4 * i + 2] = r; regionBoxes[4 * i + 3] = b;
			}
		}
	}
	if (regionCount) *regionCount = c.nStable;
	if (pointCount) *pointCount = c.np;
	free(c.nodes); free(c.ptData); free(c.ptLink); free(c.mgData); free(c.mgLink); free(bdData); free(bdLink); free(accAll); free(stackC); free(c.stableOut);
	return 0;
}
