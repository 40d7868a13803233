// Prints the texture-page plan of rr_host.hpp's planner for the texture sizes given on stdin (one per token):
//   n_slices n_nums\n nums...\n sizes...
// Used by tests/test_host_cpp.py to cross-check the C++ planner against the Python mirror (scene.plan_atlas).
#include "../openclrenderer_b200/host/rr_host.hpp"
#include <iostream>
int main() {
    std::vector<int> dims;
    for (int v; std::cin >> v;) dims.push_back(v);
    std::vector<rrhost::cl_uint> nums, sizes;
    rrhost::plan_texture_pages(dims, nums, sizes);
    std::cout << sizes.size() << " " << nums.size() << "\n";
    for (auto v : nums) std::cout << v << " ";
    std::cout << "\n";
    for (auto v : sizes) std::cout << v << " ";
    std::cout << "\n";
    return 0;
}
