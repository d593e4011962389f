// TEST INFRASTRUCTURE - see polyhedron.h in this directory.
#include <decomp_geometry/polyhedron.h>
