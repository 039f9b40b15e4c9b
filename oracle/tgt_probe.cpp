// tgt_probe.cpp -- pins the oracle's matrix/camera arithmetic against the reference's OWN tgt headers.
// TEST INFRASTRUCTURE.  Compiled by oracle/Makefile with -I/root/reference/voreen/ext (the headers are used where
// they lie; nothing from the reference is copied into this repository).  Reads commands from stdin, one per line:
//   lookat ex ey ez fx fy fz ux uy uz      -> tgt::mat4::createLookAt            (V/ext/tgt/matrix.h:1196)
//   persp fov_rad aspect near far          -> tgt::mat4::createPerspective       (V/ext/tgt/matrix.h:1230)
//   invert m0..m15                         -> tgt::mat4::invert                  (V/ext/tgt/matrix.h:1327)
//   mul a0..a15 b0..b15                    -> operator*(mat4, mat4)
//   mulv m0..m15 x y z w                   -> operator*(mat4, vec4)
//   iround f                               -> tgt::iround                        (V/ext/tgt/tgt_math.h:223)
// and prints the result as C99 hex floats (exact).
#include "tgt/matrix.h"
#include "tgt/vector.h"
#include "tgt/tgt_math.h"
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

static void print16(const tgt::mat4& m) {
    for (int i = 0; i < 16; ++i) std::printf("%a%c", m.elem[i], i == 15 ? '\n' : ' ');
}
static tgt::mat4 read16(std::istream& is) {
    tgt::mat4 m;
    for (int i = 0; i < 16; ++i) is >> m.elem[i];
    return m;
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string cmd;
        if (!(is >> cmd)) continue;
        if (cmd == "lookat") {
            tgt::vec3 e, f, u;
            is >> e.x >> e.y >> e.z >> f.x >> f.y >> f.z >> u.x >> u.y >> u.z;
            print16(tgt::mat4::createLookAt(e, f, u));
        } else if (cmd == "persp") {
            float fov, aspect, n, f;
            is >> fov >> aspect >> n >> f;
            print16(tgt::mat4::createPerspective(fov, aspect, n, f));
        } else if (cmd == "invert") {
            tgt::mat4 m = read16(is), inv;
            bool ok = m.invert(inv);
            if (!ok) std::printf("singular\n"); else print16(inv);
        } else if (cmd == "mul") {
            tgt::mat4 a = read16(is), b = read16(is);
            print16(a * b);
        } else if (cmd == "mulv") {
            tgt::mat4 a = read16(is);
            tgt::vec4 v; is >> v.x >> v.y >> v.z >> v.w;
            tgt::vec4 r = a * v;
            std::printf("%a %a %a %a\n", r.x, r.y, r.z, r.w);
        } else if (cmd == "iround") {
            float f; is >> f;
            std::printf("%d\n", tgt::iround(f));
        } else if (cmd == "deg2rad") {
            float f; is >> f;
            std::printf("%a\n", tgt::deg2rad(f));
        }
    }
    return 0;
}
