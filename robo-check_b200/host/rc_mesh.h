// rc_mesh.h -- host mesh container behind the opaque rc_mesh handle (SoA, as the reference's
// SoA loadOBJFile overload fills them: Utils_rai.cu:228-320).
#pragma once
#include <vector>

struct rc_mesh {
    std::vector<float> x, y, z;
    std::vector<int> v1, v2, v3;
};
