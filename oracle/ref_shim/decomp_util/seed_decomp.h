// TEST INFRASTRUCTURE - see ../decomp_geometry/polyhedron.h.
#include <decomp_geometry/polyhedron.h>
