// rb_scene.h — device-resident scene tables (SoA-ish flat arrays) shared by host builder and kernels.
// Layout in HBM (all read-only during a trace, a few kB .. few hundred kB, L1/L2 resident):
//   nodes[]   : one entry per PHYSICAL placed node (flattened TGeo path, DFS pre-order, 0 = top):
//               world->local rotation+translation, volume/shape/type ids, mother, BVH slice
//   bvh[]     : threaded (stackless) BVH per mother node over its daughters' world AABBs
//   shapes[], dpar[], mats[] : analytic shape parameters (+ derived constants), boolean operands
//   volumes[], borders[], indices[], graphs, th2 tables, mirrors, focals, multilayers, layers
#ifndef RB_SCENE_H
#define RB_SCENE_H

#include <stdint.h>

#include "../../include/robast_b200.h"

#if defined(__CUDACC__)
#define RB_HD __host__ __device__
#define RB_NOINLINE __noinline__
#else
#define RB_HD
#define RB_NOINLINE __attribute__((noinline))
#endif

struct DMat {  // local -> master: m = r*l + t
  double r[9];
  double t[3];
};

struct DNode {
  DMat g;            // node-local -> world
  int32_t volume;    // id into volumes[]
  int32_t shape;     // id into shapes[]
  int32_t type;      // RBG_LENS ... RBG_OTHER
  int32_t mother;    // physical id of the mother, -1 for top
  int32_t bvh_first; // slice of bvh[] over this node's daughters (bvh_count == 0: no daughters)
  int32_t bvh_count;
  int32_t overlap;   // placed with AddNodeOverlap ("MANY")
  int32_t level;     // depth in the physical tree (top = 0)
  int32_t leaf;      // RB_LEAF_*: which flat evaluator handles this node's shape (rb_device.cuh, Leaf<K>)
  float blo[3], bhi[3];  // padded world AABB of the node's shape (same box as its BVH leaf): cheap rejection of point tests
  int32_t box_first; // this node's daughters as a flat list of boxes: slice [box_first, box_first + box_count) of DScene::boxes
  int32_t box_count;
  float lc[3], lh[3];  // box of the shape in the node's own frame (centre, padded half-widths): nb_eval drops a candidate whose
                       // box the ray misses before any of its Dist* code runs — the world box of a tilted thin solid is mostly air
  int32_t pad_[1];
};

// Leaf evaluator classes.  A node's shape is a boolean tree of primitives; the generic evaluator walks it through a
// depth-bounded template recursion of non-inlined calls (Csg<DEPTH>).  Almost every shape of a real telescope is one of three
// flat patterns, which are evaluated inline instead, with the same sequence of arithmetic operations as the generic walk:
#define RB_LEAF_GENERIC 0  /* anything else: Csg<DEPTH> */
#define RB_LEAF_PRIM 1     /* a primitive */
#define RB_LEAF_BOOL2 2    /* union / intersection / subtraction of two primitives (mirror facets, cones, camera boxes) */
#define RB_LEAF_UNIONS 3   /* left-associated chain of unions whose right operands (and the innermost left one) are primitives */

struct DBvh {   // 32 B; boxes are fp32, rounded outwards and padded, traversal is fp32-conservative
  float lo[3], hi[3];
  int32_t child;     // >= 0: leaf holding this physical node id; -1: internal
  int32_t skip;      // absolute index of the next BVH entry when this subtree is done/missed (-1 = end)
};

struct DBox {   // 32 B: padded fp32 AABB of one daughter (the leaves of the mother's BVH, in BVH order)
  float lo[3], hi[3];
  int32_t child;
  int32_t pad_;
};

// operand matrix index of a boolean: -1 = identity; RB_MAT_TRANS set = pure translation (the rotation block is the unit matrix and
// is not read: "mirSphere:transZ", "boxCamera:transZ1" — most operand matrices of real telescopes)
#define RB_MAT_TRANS 0x40000000
struct DShape {
  int32_t type;
  int32_t ipar;
  int32_t left, right;
  int32_t lmat, rmat;
};

struct DScene {
  const DNode* nodes;
  const DBvh* bvh;
  const DBox* boxes;
  const DShape* shapes;
  const double* dpar;
  const DMat* mats;
  const rbg_volume* volumes;
  const rbg_border* borders;
  const rbg_graph* graphs;
  const double* gx;
  const double* gy;
  const rbg_th2* th2;
  const double* th2v;
  const rbg_index* indices;
  const rbg_mirror* mirrors;
  const rbg_focal* focals;
  const rbg_multilayer* multilayers;
  const rbg_layer* layers;
  const rbg_graph2d* graph2d;
  const int32_t* tri;
  const double* g2x;
  const double* g2y;
  const double* g2z;
  int32_t nnodes;
  int32_t top_shape;
  int32_t top_leaf;  // RB_LEAF_* of the top volume's shape
  int32_t has_many;  // some node is placed with AddNodeOverlap
  int32_t nbvh, nshapes, ndpar, nmats;  // table lengths
};

#define RB_MAX_TMM_LAYERS 16

struct DTraceParams {
  int32_t limit;
  int32_t disable_fresnel;
  uint32_t quirks;
  int32_t max_steps;     // boundary steps per launch (<=0: unbounded)
  uint64_t seed;
  uint64_t ray_id_offset;
};

#endif
