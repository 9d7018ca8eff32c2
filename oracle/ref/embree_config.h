/* Hand-written stand-in for Embree's cmake-generated kernels/config.h.
 * Mirrors the options the reference's top-level CMakeLists.txt selects
 * (/root/reference/CMakeLists.txt:20-25): defaults of ext/embree/CMakeLists.txt:119-135
 * with EMBREE_RAY_MASK off and no stat counters.  Test infrastructure only. */
/* #undef EMBREE_RAY_MASK */
/* #undef EMBREE_STAT_COUNTERS */
/* #undef EMBREE_BACKFACE_CULLING */
#define EMBREE_FILTER_FUNCTION
/* #undef EMBREE_RETURN_SUBDIV_NORMAL */
/* #undef EMBREE_IGNORE_INVALID_RAYS */
#define EMBREE_GEOMETRY_TRIANGLE
#define EMBREE_GEOMETRY_QUAD
#define EMBREE_GEOMETRY_CURVE
#define EMBREE_GEOMETRY_SUBDIVISION
#define EMBREE_GEOMETRY_USER
#define EMBREE_GEOMETRY_INSTANCE
#define EMBREE_GEOMETRY_GRID
#define EMBREE_GEOMETRY_POINT
#define EMBREE_RAY_PACKETS
#define EMBREE_CURVE_SELF_INTERSECTION_AVOIDANCE_FACTOR 2.0
#define IF_ENABLED_TRIS(x) x
#define IF_ENABLED_QUADS(x) x
#define IF_ENABLED_CURVES(x) x
#define IF_ENABLED_SUBDIV(x) x
#define IF_ENABLED_USER(x) x
#define IF_ENABLED_INSTANCE(x) x
#define IF_ENABLED_GRIDS(x) x
