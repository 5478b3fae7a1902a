// cpp_mirror_check.cpp -- the header-only C++ mirror (include/tmap_system.hpp) compiles against the
// C ABI and behaves like the reference's System on a box without a GPU (construction fails loudly)
// and, with a GPU, replays test_localmap.cpp:197-207 through the C++ surface.
#include "../include/tmap_system.hpp"
#include <cstdio>
#include <set>
using namespace traversability_mapping_b200;
int main()
{
    tmap_params p;
    tmap_default_params(&p);
    p.resolution = 0.25; p.robot_radius = 0.4; p.min_points_per_grid = 5; p.max_range = 300.0;
    try
    {
        System sys(p);
        int hits = 0;
        sys.addNewLocalMap(0, [&] { ++hits; });
        const float pts[4][4] = {{2, 0, 0, 0}, {1, 0, 0, 0}, {0, 1, 0, 0}, {2, 2, 0, 0}};
        sys.addNewKeyFrame(0, 1, 0, &pts[0][0], 4, 16);
        sys.updateKeyFrame(1, Pose::Identity());
        sys.flush();
        auto m = sys.getLocalMap(0);
        auto occ = m->occupiedCellKeys();
        std::set<std::uint64_t> got(occ.begin(), occ.end());
        const std::set<std::uint64_t> want = {Lattice::key(8, 0), Lattice::key(4, 0), Lattice::key(0, 4), Lattice::key(8, 8)};
        bool thrown = false;
        try { sys.addNewKeyFrame(0, ~0ull, 0, &pts[0][0], 4, 16); } catch (const std::runtime_error &) { thrown = true; }
        std::printf("gpu: occupied %zu callback %d negative-id-throws %d unknown-map-null %d\n", occ.size(), hits, (int)thrown,
                    (int)(sys.getLocalMap(9) == nullptr));
        return (got == want && hits == 1 && thrown && sys.getLocalMap(9) == nullptr) ? 0 : 1;
    }
    catch (const std::runtime_error &e)
    {
        std::printf("no-gpu: System() failed loudly: %s\n", e.what());
        return 42;
    }
}
