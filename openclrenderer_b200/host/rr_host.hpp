// rr_host.hpp — C++ host layer over the C ABI (include/rr.h) that keeps the reference's class and method names for the
// draw path, so that a program shaped like the reference's main.cpp (main.cpp:53-342) runs against librr_b200.so:
//
//   object_context / object_context_data   object_context.hpp:95-262, object_context.cpp
//   objects_container / object             objects_container.hpp:28-150, object.hpp
//   texture_context / texture              texture_context.hpp:27-73, texture.hpp
//   light / light_gpu                      light.hpp:13-77, light.cpp
//   engine                                 engine.hpp:100-330 (load, set_camera_pos/rot, set_light_data,
//                                          generate_realtime_shadowing, draw_bulk_objs_n, append_opencl_extra_command_line)
//   obj_load                               obj_load.cpp:181-568
//
// What is NOT here: windowing, input, networking, UI, the other renderers (out of scope). Quaternions are
// set directly (set_rot_quat): the reference's euler->quaternion conversion lives in an un-vendored library.
// Everything is in namespace rrhost so it can sit next to the reference's own headers during a migration.
// Header-only; needs zlib for PNG textures (the reference uses SFML's loader).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

extern "C" {
#include "../../include/rr.h"
}

namespace rrhost {

struct cl_float4 { float x = 0, y = 0, z = 0, w = 0; };
struct cl_float2 { float x = 0, y = 0; };
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef float cl_float;

// vertex.hpp:8-38 / triangle.hpp:8-15 — the same bytes as rr_vertex / rr_triangle
struct vertex {
    rr_vertex v{};
    void set_pos(cl_float4 p) { v.pos[0] = p.x; v.pos[1] = p.y; v.pos[2] = p.z; v.pos[3] = p.w; }
    void set_normal(cl_float4 n) { v.normal[0] = n.x; v.normal[1] = n.y; v.normal[2] = n.z; v.normal[3] = n.w; }
    void set_vt(cl_float2 t) { v.vt[0] = t.x; v.vt[1] = t.y; }
    void set_pad(cl_uint p) { v.object_id = p; }
    void set_vertex_col(uint8_t r, uint8_t g, uint8_t b, uint8_t a) { v.vertex_col = (uint32_t(r) << 24) | (uint32_t(g) << 16) | (uint32_t(b) << 8) | a; }
};
struct triangle { vertex vertices[3]; };
static_assert(sizeof(triangle) == 144, "triangle must stay 144 bytes (cl2.cl:148-151)");

[[noreturn]] inline void rr_fatal(const char* what) {      // the reference's fatal convention (ocl.h:246-251)
    std::fprintf(stderr, "%s: %s\n", what, rr_last_error());
    std::exit(4);
}

// ---------------------------------------------------------------------------------------------------------------------
// PNG (8-bit RGB / RGBA, non-interlaced) -> RGBA8. Stand-in for sf::Image::loadFromFile (texture.cpp:642).
// ---------------------------------------------------------------------------------------------------------------------
inline bool load_png_rgba(const std::string& path, std::vector<uint8_t>& out, int& w, int& h) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8)) return false;
    auto be32 = [&](size_t o) { return (uint32_t(d[o]) << 24) | (uint32_t(d[o + 1]) << 16) | (uint32_t(d[o + 2]) << 8) | d[o + 3]; };
    size_t o = 8;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat;
    while (o + 8 <= d.size()) {
        uint32_t len = be32(o);
        std::string type((const char*)&d[o + 4], 4);
        if (type == "IHDR") { w = (int)be32(o + 8); h = (int)be32(o + 12); depth = d[o + 16]; ctype = d[o + 17]; interlace = d[o + 20]; }
        else if (type == "IDAT") idat.insert(idat.end(), d.begin() + o + 8, d.begin() + o + 8 + len);
        else if (type == "IEND") break;
        o += 12 + len;
    }
    if (depth != 8 || interlace != 0 || (ctype != 6 && ctype != 2)) return false;
    const int bpp = ctype == 6 ? 4 : 3;
    const size_t stride = (size_t)w * bpp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return false;
    std::vector<uint8_t> img(stride * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* in = &raw[(stride + 1) * y];
        uint8_t* cur = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        const int ft = in[0];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0, x = in[1 + i];
            int v;
            switch (ft) {
                case 0: v = x; break;
                case 1: v = x + a; break;
                case 2: v = x + b; break;
                case 3: v = x + ((a + b) >> 1); break;
                default: { int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); v = x + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)); }
            }
            cur[i] = (uint8_t)v;
        }
    }
    out.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        out[4 * i] = img[bpp * i]; out[4 * i + 1] = img[bpp * i + 1]; out[4 * i + 2] = img[bpp * i + 2]; out[4 * i + 3] = bpp == 4 ? img[bpp * i + 3] : 255;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// textures (texture.hpp / texture_context.hpp)
// ---------------------------------------------------------------------------------------------------------------------
typedef int texture_id_t;
struct texture {
    texture_id_t id = -1;
    int gpu_id = -1;
    std::string texture_location, cache_name;
    std::vector<uint8_t> c_image;          // RGBA8 (sf::Image)
    int w = 0, h = 0;
    bool is_loaded = false, force_load = false;
    int ref_count = 1;
    void set_location(const std::string& loc) { texture_location = loc; }
    void set_texture_location(const std::string& loc) { texture_location = loc; }
    void set_create_colour(uint8_t r, uint8_t g, uint8_t b, int pw, int ph) {      // texture.cpp set_create_colour
        w = pw; h = ph; c_image.resize((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; i++) { c_image[4 * i] = r; c_image[4 * i + 1] = g; c_image[4 * i + 2] = b; c_image[4 * i + 3] = 255; }
        is_loaded = true;
    }
    void load() {
        if (is_loaded) return;
        if (!load_png_rgba(texture_location, c_image, w, h)) throw std::runtime_error("cannot load texture " + texture_location);
        is_loaded = true;
    }
    int get_largest_dimension() const { return std::max(w, h); }
    // dynamic writes into the live atlas (texture.hpp:96-99)
    void update_gpu_texture_col(const cl_float4& col, rr_ctx* dev) {                  // texture.cpp:445-463 (col in 0..255 units)
        if (id == -1 || gpu_id < 0) return;
        const float c4[4] = {col.x, col.y, col.z, col.w};
        if (rr_atlas_fill_colour(dev, (uint32_t)gpu_id, c4, (uint32_t)w, (uint32_t)h)) rr_fatal("rr_atlas_fill_colour");
    }
    void update_gpu_texture_mono(rr_ctx* dev, const uint8_t* buffer_dat, uint32_t len, int width, int height, bool flip = true) {   // texture.cpp:554-584
        if (gpu_id < 0) return;
        if (rr_atlas_upload_mono(dev, (uint32_t)gpu_id, buffer_dat, len, (uint32_t)width, (uint32_t)height, flip ? 1 : 0)) rr_fatal("rr_atlas_upload_mono");
    }
};

struct object_context;

struct texture_context_data { cl_uint mipmap_start = 0; };

struct texture_context {
    std::vector<texture*> all_textures;
    std::vector<texture_id_t> texture_id_orders;
    int gid = 0;
    cl_uint mipmap_start = 0;
    ~texture_context() { for (auto* t : all_textures) delete t; }
    texture* make_new() { texture* t = new texture; t->id = gid++; all_textures.push_back(t); return t; }
    texture* make_new_cached(const std::string& loc) {
        for (auto* t : all_textures) if (t->cache_name == loc) { t->ref_count++; return t; }
        texture* t = make_new(); t->cache_name = loc; return t;
    }
    texture* id_to_tex(int id) { for (auto* t : all_textures) if (t->id == id) return t; return nullptr; }
    int get_gpu_position_id(texture_id_t id) {
        for (size_t i = 0; i < texture_id_orders.size(); i++) if (texture_id_orders[i] == id) return (int)i;
        return -1;
    }
    texture_context_data alloc_gpu(object_context& ctx, rr_ctx* dev);      // texture_context.cpp:350-517, defined below
    // textures referenced by the active objects (+ force_load ones), ascending id: what alloc_gpu would lay out
    std::vector<texture_id_t> ids_in_use(object_context& ctx);
    std::vector<texture_id_t> built_ids;            // the set the live atlas was planned for
    std::vector<int> built_dims;
    bool built = false;
    bool should_realloc(object_context& ctx);        // texture_context.cpp should_realloc: the atlas no longer matches the scene
};

// ---------------------------------------------------------------------------------------------------------------------
// objects (object.hpp / objects_container.hpp)
// ---------------------------------------------------------------------------------------------------------------------
struct object {
    std::vector<triangle> tri_list;
    cl_float4 pos, rot_quat{0, 0, 0, 1};          // object.cpp:54
    float dynamic_scale = 1.f;
    cl_uint tid = (cl_uint)-1, rid = (cl_uint)-1, ssid = (cl_uint)-1;
    int has_bump = 0;
    float specular = 0.9f, spec_mult = 1.f, diffuse = 1.f;     // object.cpp:76-78
    int feature_flag = 0, buffer_offset = 0;
    bool isactive = false, isloaded = false;
    int object_g_id = -1, gpu_tri_start = 0, gpu_tri_end = 0;
    int object_g_id_next = -1;     // descriptor slot in the scene being built asynchronously; becomes object_g_id at flip()
    std::string object_name;
    void set_active(bool a) { isactive = a; }
    void set_pos(cl_float4 p) { pos = p; }
    void set_rot_quat(cl_float4 q) { rot_quat = q; }
    void set_dynamic_scale(float s) { dynamic_scale = s; }
    void set_feature(int flag, bool on) { feature_flag = on ? (feature_flag | flag) : (feature_flag & ~flag); }
    // object::g_flush (object.cpp:652-857): partial descriptor writes at sizeof(obj_g_descriptor)*object_g_id [+16]
    void g_flush(rr_ctx* dev) const {
        if (object_g_id < 0 || !dev) return;
        const float p[4] = {pos.x, pos.y, pos.z, 0.f}, q[4] = {rot_quat.x, rot_quat.y, rot_quat.z, rot_quat.w};
        if (rr_scene_patch_obj(dev, (uint32_t)object_g_id, (uint32_t)offsetof(rr_obj_desc, world_pos), 16, p) ||
            rr_scene_patch_obj(dev, (uint32_t)object_g_id, (uint32_t)offsetof(rr_obj_desc, world_rot_quat), 16, q) ||
            rr_scene_patch_obj(dev, (uint32_t)object_g_id, (uint32_t)offsetof(rr_obj_desc, scale), 4, &dynamic_scale))
            std::fprintf(stderr, "g_flush: %s\n", rr_last_error());
    }
};

struct objects_container {
    std::string file;
    std::vector<object> objs;
    bool isactive = false, isloaded = false, textures_are_unique = false;
    cl_float4 pos, rot_quat{0, 0, 0, 1};
    float dynamic_scale = 1.f, requested_scale = 1.f;
    int id = -1;
    object_context* parent = nullptr;
    std::function<void(objects_container*)> fp;
    void set_file(const std::string& f) { file = f; }
    void set_active(bool a) { isactive = a; for (auto& o : objs) o.set_active(a); }
    void set_pos(cl_float4 p) { pos = p; for (auto& o : objs) o.set_pos(p); }
    void set_rot_quat(cl_float4 q) { rot_quat = q; for (auto& o : objs) o.set_rot_quat(q); }
    void set_dynamic_scale(float s) { dynamic_scale = s; for (auto& o : objs) o.set_dynamic_scale(s); }
    void request_scale(float s) { requested_scale = s; }
    void set_load_func(std::function<void(objects_container*)> f) { fp = f; }
    void call_load_func(objects_container* c) { if (fp) fp(c); }
    void set_children_texture_id(cl_uint tid) { for (auto& o : objs) o.tid = tid; }              // objects_container.cpp:262
    void set_is_static(bool v) { for (auto& o : objs) o.set_feature(RR_FEATURE_IS_STATIC, v); }
    void set_two_sided(bool v) { for (auto& o : objs) o.set_feature(RR_FEATURE_TWO_SIDED, v); }
    void set_does_not_receive_dynamic_shadows(bool v) { for (auto& o : objs) o.set_feature(RR_FEATURE_NO_DYNAMIC_SHADOWS, v); }
    void set_ss_reflective(int v) { for (auto& o : objs) o.set_feature(RR_FEATURE_SS_REFLECTIVE, v != 0); }
    void set_specular(float s) { for (auto& o : objs) o.specular = s; }
    void set_spec_mult(float s) { for (auto& o : objs) o.spec_mult = s; }
    void set_diffuse(float s) { for (auto& o : objs) o.diffuse = s; }
    void set_unique_textures(bool u) { textures_are_unique = u; }
    void g_flush_objects(rr_ctx* dev) const { for (auto& o : objs) o.g_flush(dev); }
};

void obj_load(objects_container* pobj);     // obj_load.cpp:181, defined below

// ---------------------------------------------------------------------------------------------------------------------
// object_context (object_context.hpp:95-262)
// ---------------------------------------------------------------------------------------------------------------------
struct object_context_data {
    rr_ctx* dev = nullptr;
    int tri_num = 0, obj_num = 0;
    cl_uint frame_id = 0;
    cl_float4 g_clear_col;
    texture_context_data tex_gpu_ctx;
    void swap_buffers() { if (dev) rr_swap_buffers(dev); }                                        // object_context.cpp:17-25
};

struct object_context {
    std::vector<objects_container*> containers;
    texture_context tex_ctx;
    object_context_data gpu_dat;
    int cid = 0;
    ~object_context() { for (auto* c : containers) delete c; }
    objects_container* make_new() {
        objects_container* c = new objects_container;
        c->parent = this; c->id = cid++;
        c->set_load_func(obj_load);                                                                // objects_container.cpp default load func
        containers.push_back(c);
        return c;
    }
    void load_active() {                                                                           // object_context.cpp:127
        for (auto* c : containers) if (c->isactive && !c->isloaded) { c->call_load_func(c); c->set_active(true); }
    }
    object_context_data* fetch() { return &gpu_dat; }
    void set_clear_colour(const cl_float4& col) { gpu_dat.g_clear_col = col; }
    void attach(rr_ctx* dev) { gpu_dat.dev = dev; }
    // object_context::build_request / build_tick (object_context.cpp:614-640): rebuild lazily, once per frame at most
    bool request_dirty = false, rebuilding_async = false;
    void build_request() { request_dirty = true; }
    void build_tick(bool async = false) { if (request_dirty) { build(false, async); request_dirty = false; } }
    // flip_buffers (object_context.cpp:520-590): called once per frame by the render loop; swaps in an asynchronous rebuild.
    // The device waits for the uploads on its own, so the flip never blocks the host (the reference waits for an event callback).
    void flip() {
        if (!rebuilding_async) return;
        if (rr_scene_build_commit(gpu_dat.dev)) rr_fatal("rr_scene_build_commit");
        // only now do g_flush / flush_locations address the new descriptor array (until here they patched the front scene
        // with the ids it was built with); objects that moved while the upload was in flight are re-sent by the next flush
        for (auto* c : containers) for (auto& it : c->objs) it.object_g_id = c->isactive ? it.object_g_id_next : -1;
        rebuilding_async = false;
    }
    // object_context::build (object_context.cpp:646-797): textures -> descriptors (228-339) -> triangles (346-458).
    // async: the geometry goes into the device's back scene on its upload stream while frames keep rendering the current one
    // (new_gpu_dat on cqueue2 in the reference); flip() makes it current. Textures are rebuilt synchronously either way.
    void build(bool /*force*/ = false, bool async = false) {
        rr_ctx* dev = gpu_dat.dev;
        if (!dev) throw std::runtime_error("object_context::build before engine::load");
        if (rebuilding_async) flip();                                                              // object_context.cpp:651-658 ("cap")
        if (gpu_dat.tri_num == 0) async = false;                                                   // "we want there to be some valid gpu presence" (723)
        // Textures are rebuilt only when the set in use changed (texture_context::should_realloc, object_context.cpp:670-707). The
        // atlas is a single live device resource here: re-planning it moves tiles under the frames that still render the front
        // scene, so a build that changes the texture set is done synchronously as a whole (atlas and geometry switch together);
        // a build that keeps the texture set stays asynchronous and never touches the atlas.
        if (tex_ctx.should_realloc(*this)) {
            async = false;
            gpu_dat.tex_gpu_ctx = tex_ctx.alloc_gpu(*this, dev);
        }
        std::vector<rr_obj_desc> desc;
        int triangle_count = 0;
        for (auto* c : containers) {
            if (!c->isactive) continue;
            for (auto& it : c->objs) {
                rr_obj_desc d{};
                it.object_g_id_next = (int)desc.size();
                if (!async) it.object_g_id = it.object_g_id_next;
                d.tid = (cl_uint)tex_ctx.get_gpu_position_id((int)it.tid);
                d.rid = (cl_uint)tex_ctx.get_gpu_position_id((int)it.rid);
                d.ssid = (cl_uint)tex_ctx.get_gpu_position_id((int)it.ssid);
                const float p[4] = {it.pos.x, it.pos.y, it.pos.z, 0.f}, q[4] = {it.rot_quat.x, it.rot_quat.y, it.rot_quat.z, it.rot_quat.w};
                std::memcpy(d.world_pos, p, 16); std::memcpy(d.old_world_pos_1, p, 16); std::memcpy(d.old_world_pos_2, p, 16);
                std::memcpy(d.world_rot_quat, q, 16); std::memcpy(d.old_world_rot_quat_1, q, 16); std::memcpy(d.old_world_rot_quat_2, q, 16);
                d.scale = it.dynamic_scale; d.has_bump = (cl_uint)it.has_bump;
                d.specular = it.specular; d.spec_mult = it.spec_mult; d.diffuse = it.diffuse;
                d.buffer_offset = it.buffer_offset; d.feature_flag = it.feature_flag;
                it.gpu_tri_start = triangle_count;
                triangle_count += (int)it.tri_list.size();
                it.gpu_tri_end = triangle_count;
                desc.push_back(d);
            }
        }
        auto write_objs = async ? rr_scene_build_write_objs : rr_scene_write_objs;
        auto write_tris = async ? rr_scene_build_write_tris : rr_scene_write_tris;
        if ((async ? rr_scene_build_begin : rr_scene_alloc)(dev, (uint32_t)triangle_count, (uint32_t)desc.size())) rr_fatal("rr_scene_alloc");
        if (!desc.empty() && write_objs(dev, 0, (uint32_t)desc.size(), desc.data())) rr_fatal("rr_scene_write_objs");
        for (auto* c : containers) {
            if (!c->isactive) continue;
            for (auto& it : c->objs) {
                for (auto& t : it.tri_list) t.vertices[0].set_pad((cl_uint)it.object_g_id_next);  // object_context.cpp:427 / fill_ids
                if (!it.tri_list.empty() &&
                    write_tris(dev, (uint32_t)it.gpu_tri_start, (uint32_t)it.tri_list.size(), (const rr_triangle*)it.tri_list.data()))
                    rr_fatal("rr_scene_write_tris");
            }
        }
        gpu_dat.tri_num = triangle_count; gpu_dat.obj_num = (int)desc.size();
        rebuilding_async = async;
    }
    void flush_locations() { for (auto* c : containers) if (c->isactive) c->g_flush_objects(gpu_dat.dev); }   // object_context.cpp:819
};

// The page planner of texture_context::alloc_gpu as a pure function (texture_context.cpp:94-261): `dims` = largest dimension
// of every texture in gpu-id order. nums[i] = slice << 16 | index for the base levels, then nums[n + 4*i + level] for the
// four mips of texture i; sizes[slice] = tile size of the slice. Pages are keyed by size ascending (std::map) and indices
// count down from the page's population. Cross-checked against the Python mirror (scene.plan_atlas) on thousands of
// textures by tests/test_host_cpp.py.
inline void plan_texture_pages(const std::vector<int>& dims, std::vector<cl_uint>& nums, std::vector<cl_uint>& sizes) {
    const int MIPS = 4, MAXSZ = 2048;
    std::map<size_t, int> size_to_numbers;
    for (int s : dims) {
        size_to_numbers[s]++;
        for (int j = 0; j < MIPS; j++) size_to_numbers[s / (1 << (j + 1))]++;
    }
    struct page { int size, n; };
    std::vector<page> pages;
    for (auto& kv : size_to_numbers) {
        if (kv.first == 0) throw std::runtime_error("texture too small for 4 mip levels");
        int per = (MAXSZ / (int)kv.first) * (MAXSZ / (int)kv.first), rem = kv.second;
        while (rem >= per) { pages.push_back({(int)kv.first, per}); rem -= per; }
        if (rem > 0) pages.push_back({(int)kv.first, rem});
    }
    std::vector<page> free_pages = pages;
    // first page of every size, so that thousands of textures do not rescan the page list for each tile
    std::map<int, size_t> first_of_size;
    for (size_t sl = free_pages.size(); sl-- > 0;) first_of_size[free_pages[sl].size] = sl;
    auto take = [&](int size) -> cl_uint {
        auto it = first_of_size.find(size);
        if (it != first_of_size.end())
            for (size_t sl = it->second; sl < free_pages.size() && free_pages[sl].size == size; sl++)
                if (free_pages[sl].n > 0) { free_pages[sl].n--; it->second = sl; return (cl_uint)((sl << 16) | (cl_uint)free_pages[sl].n); }
        throw std::runtime_error("could not find a free texture page");
    };
    nums.clear(); sizes.clear();
    for (int s : dims) nums.push_back(take(s));
    for (int s : dims) for (int j = 0; j < MIPS; j++) nums.push_back(take(s / (1 << (j + 1))));
    for (auto& p : pages) sizes.push_back((cl_uint)p.size);
}

// texture_context::alloc_gpu: page planner (texture_context.cpp:94-261) + uploads (texture.cpp:323-358, 465-493)
inline std::vector<texture_id_t> texture_context::ids_in_use(object_context& ctx) {
    std::set<texture_id_t> in_use;
    for (auto* c : ctx.containers) {
        if (!c->isactive) continue;
        for (auto& o : c->objs) {
            if ((int)o.tid != -1) in_use.insert((int)o.tid);
            if ((int)o.rid != -1) in_use.insert((int)o.rid);
            if ((int)o.ssid != -1) in_use.insert((int)o.ssid);
        }
    }
    for (auto* t : all_textures) if (t->force_load) in_use.insert(t->id);
    return std::vector<texture_id_t>(in_use.begin(), in_use.end());
}

inline bool texture_context::should_realloc(object_context& ctx) {
    if (!built) return true;
    const std::vector<texture_id_t> ids = ids_in_use(ctx);
    if (ids != built_ids) return true;
    for (size_t i = 0; i < ids.size(); i++) {
        texture* t = id_to_tex(ids[i]);
        t->load();
        if (t->get_largest_dimension() != built_dims[i]) return true;
    }
    return false;
}

inline texture_context_data texture_context::alloc_gpu(object_context& ctx, rr_ctx* dev) {
    const std::vector<texture_id_t> in_use = ids_in_use(ctx);
    texture_id_orders.assign(in_use.begin(), in_use.end());
    mipmap_start = (cl_uint)in_use.size();
    for (auto id : in_use) id_to_tex(id)->load();
    std::vector<int> dims;
    for (auto id : in_use) dims.push_back(id_to_tex(id)->get_largest_dimension());
    std::vector<cl_uint> nums, sizes;
    plan_texture_pages(dims, nums, sizes);
    if (rr_atlas_alloc(dev, (uint32_t)sizes.size(), nums.data(), (uint32_t)nums.size(), sizes.data(), (uint32_t)sizes.size(), mipmap_start)) rr_fatal("rr_atlas_alloc");
    // texture_context.cpp:478-517 uploads the textures one by one (write + kernel + 4 mip kernels each); here the whole set goes
    // down in one staged copy and one launch per phase
    std::vector<uint32_t> gpu_ids, ws, hs;
    std::vector<const uint8_t*> images;
    int c = 0;
    for (auto id : texture_id_orders) {
        texture* t = id_to_tex(id);
        t->gpu_id = c++;
        gpu_ids.push_back((uint32_t)t->gpu_id); images.push_back(t->c_image.data()); ws.push_back((uint32_t)t->w); hs.push_back((uint32_t)t->h);
    }
    if (rr_atlas_upload_batch(dev, (uint32_t)gpu_ids.size(), gpu_ids.data(), images.data(), ws.data(), hs.data(), 1)) rr_fatal("rr_atlas_upload_batch");
    built_ids = in_use; built_dims = dims; built = true;
    texture_context_data d; d.mipmap_start = mipmap_start;
    return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// obj_load (obj_load.cpp:181-568): triangulated OBJ, explicit vt / vn, one object per usemtl, map_Kd from the .mtl
// ---------------------------------------------------------------------------------------------------------------------
inline void obj_load(objects_container* pobj) {
    const std::string filename = pobj->file;
    const std::string mtlname = filename.substr(0, filename.find_last_of('.')) + ".mtl";
    const size_t lslash = filename.find_last_of('/');
    const std::string dir = lslash == std::string::npos ? "." : filename.substr(0, lslash);
    std::ifstream file(filename), mtlfile(mtlname);
    if (!file.is_open()) { std::fprintf(stderr, "%s could not be opened in obj_load\n", filename.c_str()); return; }
    std::vector<std::string> mtl;
    for (std::string ln; std::getline(mtlfile, ln);) { if (!ln.empty() && ln.back() == '\r') ln.pop_back(); mtl.push_back(ln); }
    std::vector<cl_float4> vl, vnl;
    std::vector<cl_float2> vtl;
    struct idx { int v[3], vt[3], vn[3]; };
    std::vector<idx> fl;
    std::vector<int> usemtl_pos;
    std::vector<std::string> usemtl_name;
    for (std::string ln; std::getline(file, ln);) {
        if (!ln.empty() && ln.back() == '\r') ln.pop_back();
        if (ln.size() < 2) continue;
        if (ln[0] == 'f' && ln[1] == ' ') {
            idx f{};
            int start = 2;
            for (int i = 0; i < 3; i++) {                                   // decompose_face, obj_load.cpp:142-158
                size_t s1 = ln.find('/', start), s2 = ln.find('/', s1 + 1), s3 = ln.find(' ', s2 + 1);
                f.v[i] = std::atoi(ln.c_str() + start) - 1;
                f.vt[i] = std::atoi(ln.c_str() + s1 + 1) - 1;
                f.vn[i] = std::atoi(ln.c_str() + s2 + 1) - 1;
                start = (int)s3 + 1;
            }
            fl.push_back(f);
        } else if (ln[0] == 'v' && ln[1] == ' ') { cl_float4 t; std::sscanf(ln.c_str() + 2, "%f %f %f", &t.x, &t.y, &t.z); vl.push_back(t); }
        else if (ln.compare(0, 3, "vt ") == 0) { cl_float2 t; std::sscanf(ln.c_str() + 3, "%f %f", &t.x, &t.y); vtl.push_back(t); }
        else if (ln.compare(0, 3, "vn ") == 0) { cl_float4 t; std::sscanf(ln.c_str() + 3, "%f %f %f", &t.x, &t.y, &t.z); vnl.push_back(t); }
        else if (ln.compare(0, 3, "use") == 0) { usemtl_pos.push_back((int)fl.size()); usemtl_name.push_back(ln.substr(ln.find_last_of(' ') + 1)); }
    }
    std::vector<triangle> tris(fl.size());
    const float rs = pobj->requested_scale;
    for (size_t i = 0; i < fl.size(); i++)
        for (int j = 0; j < 3; j++) {
            cl_float4 v = vl[fl[i].v[j]];
            tris[i].vertices[j].set_pos({v.x * rs, v.y * rs, v.z * rs, 0});
            tris[i].vertices[j].set_vt(vtl[fl[i].vt[j]]);
            tris[i].vertices[j].set_normal(vnl[fl[i].vn[j]]);
        }
    usemtl_pos.push_back((int)tris.size());
    texture_context* tex_ctx = &pobj->parent->tex_ctx;
    for (size_t i = 0; i + 1 < usemtl_pos.size(); i++) {
        std::string texture_name;                                            // retrieve_diffuse_new, obj_load.cpp:21-47
        bool found = false;
        for (auto& ln : mtl) {
            if (ln.compare(0, 7, "newmtl ") == 0) found = ln.substr(ln.find_last_of(' ') + 1) == usemtl_name[i];
            else if (found && ln.compare(0, 7, "map_Kd ") == 0) { texture_name = ln.substr(ln.find_last_of(' ') + 1); break; }
        }
        texture* tex;
        const std::string full = dir + "/" + texture_name;
        if (!texture_name.empty() && std::ifstream(full).good()) {
            tex = pobj->textures_are_unique ? tex_ctx->make_new() : tex_ctx->make_new_cached(full);
            tex->set_location(full);
        } else {
            tex = tex_ctx->make_new();
            tex->set_create_colour(255, 0, 255, 32, 32);                     // obj_load.cpp:497-500
        }
        object obj;
        obj.tri_list.assign(tris.begin() + usemtl_pos[i], tris.begin() + usemtl_pos[i + 1]);
        obj.tid = (cl_uint)tex->id;
        obj.pos = pobj->pos; obj.rot_quat = pobj->rot_quat; obj.dynamic_scale = pobj->dynamic_scale;
        obj.isloaded = true;
        pobj->objs.push_back(obj);
    }
    pobj->requested_scale = 1.f;
    pobj->isloaded = true;
}

// ---------------------------------------------------------------------------------------------------------------------
// lights (light.hpp:13-77)
// ---------------------------------------------------------------------------------------------------------------------
struct light_gpu { int n_lights = 0; };

struct light {
    cl_float4 pos, col{1, 1, 1, 0};
    cl_uint shadow = 0;
    cl_float brightness = 1.f, radius = 100000.f, diffuse = 1.f, godray_intensity = 0.f;
    cl_int is_static = 0;
    void set_pos(cl_float4 p) { pos = p; }
    void set_col(cl_float4 c) { col = c; }
    void set_shadow_casting(cl_uint s) { shadow = s; }
    void set_brightness(cl_float b) { brightness = b; }
    void set_radius(cl_float r) { radius = r; }
    void set_diffuse(cl_float d) { diffuse = d; }
    void set_godray_intensity(cl_float g) { godray_intensity = g; }
    void set_is_static(bool s) { is_static = s; }
    static std::vector<light*>& lightlist() { static std::vector<light*> l; return l; }
    static std::vector<cl_uint>& active() { static std::vector<cl_uint> a; return a; }
    static bool& static_lights_are_dirty() { static bool d = true; return d; }
    static light* add_light(const light* l) { light* n = new light(*l); lightlist().push_back(n); active().push_back(1); static_lights_are_dirty() = true; return n; }
    static void remove_light(light* l) {
        auto& ll = lightlist();
        for (size_t i = 0; i < ll.size(); i++) if (ll[i] == l) { delete l; ll.erase(ll.begin() + i); active().erase(active().begin() + i); break; }
    }
    // light::build (light.cpp:145-276): active lights packed in list order; cubemap slabs are (re)allocated by the device layer
    static light_gpu build(rr_ctx* dev) {
        std::vector<rr_light> packed;
        for (size_t i = 0; i < lightlist().size(); i++) {
            if (!active()[i]) continue;
            const light& l = *lightlist()[i];
            rr_light r{};
            r.pos[0] = l.pos.x; r.pos[1] = l.pos.y; r.pos[2] = l.pos.z; r.pos[3] = l.pos.w;
            r.col[0] = l.col.x; r.col[1] = l.col.y; r.col[2] = l.col.z; r.col[3] = l.col.w;
            r.shadow = l.shadow; r.brightness = l.brightness; r.radius = l.radius; r.diffuse = l.diffuse;
            r.godray_intensity = l.godray_intensity; r.is_static = l.is_static;
            packed.push_back(r);
        }
        if (rr_lights_write(dev, packed.data(), (uint32_t)packed.size())) rr_fatal("rr_lights_write");
        light_gpu g; g.n_lights = (int)packed.size();
        return g;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// engine (engine.hpp:100-330) — the draw path only
// ---------------------------------------------------------------------------------------------------------------------
struct engine {
    rr_ctx* dev = nullptr;
    rr_config cfg{};
    cl_uint width = 0, height = 0, depth = 0;
    cl_float4 c_pos, c_rot;
    light_gpu* light_data = nullptr;
    std::vector<std::string> extra;
    engine() { rr_default_config(&cfg); }
    ~engine() { if (dev) rr_destroy(dev); }
    // the reference passes its configuration as OpenCL -D macros (main.cpp:80-85); the same strings are understood here
    void append_opencl_extra_command_line(const std::string& str) {
        extra.push_back(str);
        auto val = [&](const char* key, float& out) { size_t p = str.find(key); if (p != std::string::npos) out = std::strtof(str.c_str() + p + std::strlen(key), nullptr); };
        float f;
        f = (float)cfg.depth_icutoff; val("depth_icutoff=", f); cfg.depth_icutoff = (int)f;
        val("AMBIENT=", cfg.ambient); val("SSAO_RAD=", cfg.ssao_rad); val("SSAO_DIV=", cfg.ssao_div); val("MIP_BIAS=", cfg.mip_bias);
        val("SHADOWBIAS=", cfg.shadow_bias); val("SHADOWEXP=", cfg.shadow_exp);
        if (str.find("TEST_LINEAR") != std::string::npos) cfg.test_linear = 1;
        if (str.find("NO_SSAO") != std::string::npos) cfg.no_ssao = 1;
    }
    void load(cl_uint pwidth, cl_uint pheight, cl_uint pdepth, const std::string& /*name*/, const std::string& /*kernel file*/, bool /*only_3d*/ = false) {
        width = pwidth; height = pheight; depth = pdepth;
        cfg.width = (int)width; cfg.height = (int)height;
        dev = rr_create(&cfg);
        if (!dev) rr_fatal("rr_create");
    }
    void set_camera_pos(cl_float4 p) { c_pos = p; }
    void set_camera_rot(cl_float4 r) { c_rot = r; }
    void set_light_data(light_gpu& ld) { light_data = &ld; }
    void generate_realtime_shadowing(object_context_data&) {
        if (rr_frame_shadows(dev, light::static_lights_are_dirty())) std::fprintf(stderr, "generate_realtime_shadowing: %s\n", rr_last_error());
        light::static_lights_are_dirty() = false;
    }
    void draw_bulk_objs_n(object_context_data& dat) {
        const float p[4] = {c_pos.x, c_pos.y, c_pos.z, 0}, r[4] = {c_rot.x, c_rot.y, c_rot.z, 0};
        const float cl[4] = {dat.g_clear_col.x, dat.g_clear_col.y, dat.g_clear_col.z, dat.g_clear_col.w};
        if (rr_frame_draw(dev, p, r, cl)) std::fprintf(stderr, "draw_bulk_objs_n: %s\n", rr_last_error());
        dat.frame_id++;
    }
    // headless replacement of blit_to_screen / render_block / flip (engine.cpp:3398-3820)
    void blit_to_host(std::vector<uint8_t>& rgba) { rgba.resize((size_t)width * height * 4); if (rr_read_rgba8(dev, rgba.data())) rr_fatal("rr_read_rgba8"); }
};

}  // namespace rrhost
