// main_headless.cpp — the reference's main.cpp (main.cpp:53-342), headless, over librr_b200.so through the host layer that
// keeps the reference's class names (openclrenderer_b200/host/rr_host.hpp).
//
//   main_headless <model.obj> <out_prefix> [w h scale cam_x cam_y cam_z rot_x light_x light_y light_z shadow frames]
// writes <out_prefix>.depth / .ids / .rgba (raw little-endian buffers) of the last frame.
#include <cstdio>
#include <cstdlib>

#include "../openclrenderer_b200/host/rr_host.hpp"

using namespace rrhost;

static void dump(const std::string& path, const void* p, size_t n) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f || std::fwrite(p, 1, n, f) != n) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); std::exit(3); }
    std::fclose(f);
}

int main(int argc, char* argv[]) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s model.obj out_prefix [w h scale cx cy cz rx lx ly lz shadow frames]\n", argv[0]); return 2; }
    auto arg = [&](int i, double d) { return argc > i ? std::atof(argv[i]) : d; };
    const int w = (int)arg(3, 800), h = (int)arg(4, 600);
    const float scale = (float)arg(5, 100);

    object_context context;
    objects_container* model = context.make_new();
    model->set_file(argv[1]);
    model->set_active(true);

    engine window;
    window.append_opencl_extra_command_line("-D depth_icutoff=20");          // main.cpp:80-85
    window.append_opencl_extra_command_line("-D AMBIENT=0.2f");
    window.append_opencl_extra_command_line("-D SSAO_RAD=2.f");
    window.append_opencl_extra_command_line("-D TEST_LINEAR");
    window.load(w, h, 1000, "turtles", "cl2.cl", true);
    context.attach(window.dev);

    window.set_camera_pos({(float)arg(6, 0), (float)arg(7, 150), (float)arg(8, -400), 0});
    window.set_camera_rot({(float)arg(9, 0.3), 0, 0, 0});

    context.load_active();
    model->set_dynamic_scale(scale);
    context.build(true);

    light l;
    l.set_col({1.f, 1.f, 1.f, 0.f});
    l.set_shadow_casting((cl_uint)arg(13, 0));
    l.set_brightness(1.f);
    l.set_radius(20000.f);
    l.set_pos({(float)arg(10, -200), (float)arg(11, 300), (float)arg(12, -300), 0});
    light::add_light(&l);
    light_gpu light_data = light::build(window.dev);
    window.set_light_data(light_data);

    const int frames = (int)arg(14, 2);
    for (int i = 0; i < frames; i++) {                                        // main.cpp:243-291
        if (i) context.fetch()->swap_buffers();
        window.generate_realtime_shadowing(*context.fetch());
        window.draw_bulk_objs_n(*context.fetch());
    }
    if (rr_sync(window.dev)) rr_fatal("rr_sync");

    const size_t P = (size_t)w * h;
    std::vector<uint32_t> depth(P), ids(P);
    std::vector<uint8_t> rgba;
    if (rr_read_depth(window.dev, depth.data()) || rr_read_ids(window.dev, ids.data())) rr_fatal("read");
    window.blit_to_host(rgba);
    const std::string out = argv[2];
    dump(out + ".depth", depth.data(), P * 4);
    dump(out + ".ids", ids.data(), P * 4);
    dump(out + ".rgba", rgba.data(), P * 4);
    size_t covered = 0;
    for (auto d : depth) covered += d != 0xFFFFFFFFu;
    rr_timings t;
    rr_get_timings(window.dev, &t);
    std::printf("tris %d objs %d covered %zu fragments %u frame %.3f ms\n", context.fetch()->tri_num, context.fetch()->obj_num, covered, t.n_fragments, t.frame_ms);
    return 0;
}
