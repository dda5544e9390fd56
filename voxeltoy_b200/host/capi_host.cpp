// capi_host.cpp -- flat C wrappers over the host classes (vt_host.h) so that Python (ctypes) tests and the
// bench can drive Renderer / loaders / tools exactly the way the reference's UI drives them. Thin: one call each.
#include "vt_host.h"

#include <cstring>

using namespace vtm;

namespace {
struct VoxResult { std::vector<int32_t> grid, emissive; std::vector<float> materials; V3i res; std::string error; };
struct ObjResult { std::vector<float> verts; std::vector<unsigned int> idx; };
struct CdfResult { std::vector<float> cdfU, cdfV; unsigned int w, h; float integral; };
struct CollectLogger : public Logger {
    std::string all;
    void operator()(const std::string& msg) override { all += msg; all += "\n"; }
};
struct RendererBox { Renderer r; CollectLogger log; std::vector<Material::SerializedData> mats; };
}

extern "C" {

// ---- Renderer ---------------------------------------------------------------------------------------------------
void* vth_renderer_create(void) { RendererBox* b = new RendererBox(); b->r.setLogger(&b->log); return b; }
void vth_renderer_destroy(void* h) { delete static_cast<RendererBox*>(h); }
#define R(h) (static_cast<RendererBox*>(h)->r)
int vth_renderer_initialize(void* h, int device) { R(h).initializeOnDevice(device); return R(h).context() ? 0 : -1; }
void vth_renderer_resize_frame(void* h, int w, int hgt, int vx, int vy, int vw, int vh) { R(h).resizeFrame(w, hgt, vx, vy, vw, vh); }
int vth_renderer_render(void* h) { return (int)R(h).render(); }
int vth_renderer_render_passes(void* h, int n) { return (int)R(h).renderPasses(n); }
void vth_renderer_reload_shaders(void* h, const char* p) { R(h).reloadShaders(p ? p : ""); }
void vth_renderer_load_vox_file(void* h, const char* path) { R(h).loadVoxFile(path); }
void vth_renderer_load_mesh(void* h, const char* path, int resolution) { if (resolution > 0) R(h).loadMeshAtResolution(path, resolution); else R(h).loadMesh(path); }
void vth_renderer_set_voxel_data(void* h, int X, int Y, int Z, const int32_t* grid, const float* mats, size_t nm, const int32_t* em, size_t ne)
{
    std::vector<int32_t> g(grid, grid + (size_t)X * Y * Z), e(em, em + (em ? ne : 0));
    std::vector<float> m(mats, mats + (mats ? nm : 0));
    R(h).setVoxelData(V3i(X, Y, Z), g, m, e);
}
void vth_renderer_save_image(void* h, const char* path) { R(h).saveImage(path); }
int vth_renderer_read_average(void* h, float* out) { return R(h).readAverage(out) ? 0 : -1; }
void vth_renderer_reset_render(void* h) { R(h).resetRender(); }
int vth_renderer_on_mouse_move(void* h, int dx, int dy, int buttons) { return R(h).onMouseMove(dx, dy, buttons) ? 1 : 0; }
int vth_renderer_on_key_press(void* h, int key) { return R(h).onKeyPress(key) ? 1 : 0; }
void vth_renderer_request_action(void* h, float x, float y, float dx, float dy, int action, int restart)
{
    R(h).requestAction(x, y, dx, dy, (Action::PICKING_ACTION)action, restart != 0);
}
void vth_renderer_update_render_settings(void* h) { R(h).updateRenderSettings(); }
const char* vth_renderer_status(void* h) { return R(h).getStatus().c_str(); }
const char* vth_renderer_log(void* h) { return static_cast<RendererBox*>(h)->log.all.c_str(); }
void* vth_renderer_context(void* h) { return R(h).context(); }
int vth_renderer_number_samples(void* h) { return R(h).numberSamples(); }
void vth_renderer_set_integrator(void* h, int i) { R(h).setIntegrator((Renderer::Integrator)i); }
void vth_renderer_set_partition(void* h, int mode, int rank, int world) { R(h).setPartition(mode, rank, world); }
void vth_renderer_camera_matrices(void* h, float* imv, float* pm, float* ipm) { R(h).cameraMatrices(imv, pm, ipm); }
void vth_renderer_volume_info(void* h, int* res, float* bmin, float* bmax)
{
    const V3i r = R(h).volumeResolution(); const Box3f b = R(h).volumeBounds();
    res[0] = r.x; res[1] = r.y; res[2] = r.z;
    for (int i = 0; i < 3; ++i) { bmin[i] = b.min[i]; bmax[i] = b.max[i]; }
}
// RenderSettings fields (renderer.renderSettings().m_* in C++)
void vth_settings_set(void* h, int maxBounces, int maxSamples, float wireOpacity, float wireThickness,
                      const float* bgTop, const float* bgBottom, const char* bgImage, int rotationDegrees)
{
    RenderSettings& s = R(h).renderSettings();
    if (maxBounces >= 0) s.m_pathtracerMaxNumBounces = maxBounces;
    if (maxSamples > 0) s.m_pathtracerMaxSamples = maxSamples;
    if (wireOpacity >= 0) s.m_wireframeOpacity = wireOpacity;
    if (wireThickness >= 0) s.m_wireframeThickness = wireThickness;
    if (bgTop) s.m_backgroundColor[0] = V3f(bgTop[0], bgTop[1], bgTop[2]);
    if (bgBottom) s.m_backgroundColor[1] = V3f(bgBottom[0], bgBottom[1], bgBottom[2]);
    if (bgImage) s.m_backgroundImage = bgImage;
    s.m_backgroundRotationDegrees = rotationDegrees;
}
// Camera (renderer.camera().* in C++)
void vth_camera_set_lens_model(void* h, int m) { R(h).camera().setLensModel((CameraParameters::CameraLensModel)m); }
void vth_camera_set_fstop(void* h, float f) { R(h).camera().setFStop(f); }
void vth_camera_set_focal_length(void* h, float f) { R(h).camera().setFocalLength(f); }
void vth_camera_set_lens_radius(void* h, float f) { R(h).camera().setLensRadius(f); }
void vth_camera_set_controller(void* h, int mode) { R(h).camera().setCameraController((Camera::CameraControllerMode)mode); }
void vth_camera_look_at(void* h, float x, float y, float z) { R(h).camera().controller().lookAt(V3f(x, y, z)); }
void vth_camera_set_distance(void* h, float d) { R(h).camera().controller().setDistanceFromTarget(d); }
void vth_camera_orbit(void* h, float theta, float phi) { R(h).camera().controller().orbitAroundTarget(theta, phi); }
void vth_camera_get(void* h, float* eye, float* target, float* scalars /* fovY, focalLength, lensRadius, near, far, filmW, filmH, lensModel */)
{
    const CameraParameters& p = R(h).camera().parameters();
    for (int i = 0; i < 3; ++i) { eye[i] = p.eye()[i]; target[i] = p.target()[i]; }
    scalars[0] = p.fovY(); scalars[1] = p.focalLength(); scalars[2] = p.lensRadius(); scalars[3] = p.nearDistance();
    scalars[4] = p.farDistance(); scalars[5] = p.filmSize().x; scalars[6] = p.filmSize().y; scalars[7] = (float)p.lensModel();
}
// materials
int vth_renderer_get_materials(void* h, int maxN, int* types, int* offsets)
{
    RendererBox* b = static_cast<RendererBox*>(h);
    b->mats = b->r.getMaterials();
    int n = 0;
    for (size_t i = 0; i < b->mats.size() && n < maxN; ++i, ++n) {
        const std::string& nm = b->mats[i].m_propertyName;
        types[n] = nm[0] == 'L' ? 0 : (nm[0] == 'M' ? 1 : 2);
        offsets[n] = (int)b->mats[i].m_dataOffset;
    }
    return (int)b->mats.size();
}
void vth_renderer_update_material_color(void* h, unsigned int off, const float* rgb) { R(h).updateMaterialColor(off, rgb); }
void vth_renderer_update_material_value(void* h, unsigned int off, float v) { R(h).updateMaterialValue(off, v); }

// ---- tools ------------------------------------------------------------------------------------------------------------
void* vth_tool_create(void* renderer, int kind) { return kind == 0 ? (Tool*)new ToolAddRemoveVoxel(R(renderer)) : (Tool*)new ToolFocalDistance(R(renderer)); }
void vth_tool_destroy(void* t) { delete static_cast<Tool*>(t); }
int vth_tool_mouse(void* t, int press, int x, int y, int buttons, int modifiers, int w, int hgt)
{
    MouseEvent e = { x, y, buttons, modifiers }; WidgetSize s = { w, hgt };
    Tool* tool = static_cast<Tool*>(t);
    return (press ? tool->mousePressEvent(&e, s) : tool->mouseMoveEvent(&e, s)) ? 1 : 0;
}

// ---- headless widget (ui/glwidget.cpp without Qt) ----------------------------------------------------------------------
void* vth_widget_create(void* renderer) { return new HeadlessWidget(R(renderer)); }
void vth_widget_destroy(void* w) { delete static_cast<HeadlessWidget*>(w); }
int vth_widget_run_script(void* w, const char* script, char* error, int errorSize)
{
    std::string err;
    const bool ok = static_cast<HeadlessWidget*>(w)->runScript(script ? script : "", err);
    if (error && errorSize > 0) { strncpy(error, err.c_str(), (size_t)errorSize - 1); error[errorSize - 1] = 0; }
    return ok ? 0 : -1;
}
int vth_widget_pump(void* w, int maxPaints) { return static_cast<HeadlessWidget*>(w)->pump(maxPaints); }
int vth_widget_update_pending(void* w) { return static_cast<HeadlessWidget*>(w)->updatePending() ? 1 : 0; }
unsigned long vth_widget_paints(void* w) { return static_cast<HeadlessWidget*>(w)->paints(); }
void vth_renderer_resolution(void* r, int* w, int* h) { *w = R(r).renderSettings().m_imageResolution.x; *h = R(r).renderSettings().m_imageResolution.y; }
void vth_widget_size(void* w, int* width, int* height) { *width = static_cast<HeadlessWidget*>(w)->width(); *height = static_cast<HeadlessWidget*>(w)->height(); }

// ---- loaders -------------------------------------------------------------------------------------------------------------
// palette rules: n records of (colour index, material type, emission r g b, roughness) as 6 floats each
static std::vector<MagicaVoxelLoader::PaletteRule> palette_rules(const float* rules, int n)
{
    std::vector<MagicaVoxelLoader::PaletteRule> out;
    for (int i = 0; rules && i < n; ++i) {
        MagicaVoxelLoader::PaletteRule r;
        r.colorIndex = (int)rules[6 * i]; r.type = (Material::MaterialType)(int)rules[6 * i + 1];
        r.emission = V3f(rules[6 * i + 2], rules[6 * i + 3], rules[6 * i + 4]); r.roughness = rules[6 * i + 5];
        out.push_back(r);
    }
    return out;
}
void vth_renderer_set_vox_palette_rules(void* h, const float* rules, int n) { R(h).setVoxPaletteRules(palette_rules(rules, n)); }
void* vth_vox_load_rules(const char* path, const float* rules, int n)
{
    VoxResult* r = new VoxResult();
    MagicaVoxelLoader loader;
    loader.setPaletteRules(palette_rules(rules, n));
    if (!loader.load(path, r->grid, r->materials, r->emissive, r->res)) { r->error = loader.m_error; r->res = V3i(0); }
    return r;
}
void* vth_vox_load(const char* path)
{
    VoxResult* r = new VoxResult();
    MagicaVoxelLoader loader;
    if (!loader.load(path, r->grid, r->materials, r->emissive, r->res)) { r->error = loader.m_error; r->res = V3i(0); }
    return r;
}
const char* vth_vox_error(void* h) { return static_cast<VoxResult*>(h)->error.c_str(); }
void vth_vox_dims(void* h, int* res, size_t* nm, size_t* ne)
{
    VoxResult* r = static_cast<VoxResult*>(h);
    res[0] = r->res.x; res[1] = r->res.y; res[2] = r->res.z; *nm = r->materials.size(); *ne = r->emissive.size();
}
void vth_vox_copy(void* h, int32_t* grid, float* mats, int32_t* em)
{
    VoxResult* r = static_cast<VoxResult*>(h);
    if (!r->grid.empty()) memcpy(grid, &r->grid[0], r->grid.size() * 4);
    if (!r->materials.empty()) memcpy(mats, &r->materials[0], r->materials.size() * 4);
    if (!r->emissive.empty()) memcpy(em, &r->emissive[0], r->emissive.size() * 4);
}
void vth_vox_free(void* h) { delete static_cast<VoxResult*>(h); }

void* vth_obj_load(const char* path) { ObjResult* r = new ObjResult(); MeshLoader::loadFromOBJ(path, r->verts, r->idx); return r; }
void vth_obj_dims(void* h, size_t* nv, size_t* ni) { ObjResult* r = static_cast<ObjResult*>(h); *nv = r->verts.size() / 3; *ni = r->idx.size(); }
void vth_obj_copy(void* h, float* v, unsigned int* i)
{
    ObjResult* r = static_cast<ObjResult*>(h);
    if (!r->verts.empty()) memcpy(v, &r->verts[0], r->verts.size() * 4);
    if (!r->idx.empty()) memcpy(i, &r->idx[0], r->idx.size() * 4);
}
void vth_obj_free(void* h) { delete static_cast<ObjResult*>(h); }
void vth_compute_mesh_transform(const float* bmin, const float* bmax, const int* res, float* out16)
{
    const M44f m = computeMeshTransform(Box3f(V3f(bmin[0], bmin[1], bmin[2]), V3f(bmax[0], bmax[1], bmax[2])), V3i(res[0], res[1], res[2]));
    memcpy(out16, m.x, 64);
}
void vth_prune_emissive(const int32_t* grid, int X, int Y, int Z, int32_t* em, size_t* n)
{
    Renderer r;      // uninitialised renderer: pruneInteriorEmissiveVoxels is pure host code
    std::vector<int32_t> g(grid, grid + (size_t)X * Y * Z), e(em, em + *n);
    V3i res(X, Y, Z);
    r.pruneInteriorEmissiveVoxels(g, res, e);
    if (!e.empty()) memcpy(em, &e[0], e.size() * 4);
    *n = e.size();
}

// ---- environment --------------------------------------------------------------------------------------------------------------
void* vth_cdf_build(const float* rgb, unsigned int w, unsigned int h)
{
    CdfResult* r = new CdfResult();
    if (!calculateCDF(rgb, w, h, r->cdfU, r->w, r->h, r->cdfV, r->integral)) { r->w = r->h = 0; }
    return r;
}
void vth_cdf_dims(void* hd, unsigned int* w, unsigned int* h, float* integral) { CdfResult* r = static_cast<CdfResult*>(hd); *w = r->w; *h = r->h; *integral = r->integral; }
void vth_cdf_copy(void* hd, float* u, float* v)
{
    CdfResult* r = static_cast<CdfResult*>(hd);
    if (!r->cdfU.empty()) memcpy(u, &r->cdfU[0], r->cdfU.size() * 4);
    if (!r->cdfV.empty()) memcpy(v, &r->cdfV[0], r->cdfV.size() * 4);
}
void vth_cdf_free(void* hd) { delete static_cast<CdfResult*>(hd); }
int vth_write_pfm(const char* path, const float* rgb, unsigned int w, unsigned int h) { return writePFM(path, rgb, w, h) ? 0 : -1; }
int vth_write_png(const char* path, const unsigned char* rgba8, unsigned int w, unsigned int h) { return writePNG(path, rgba8, w, h) ? 0 : -1; }
int vth_write_exr(const char* path, const float* pixels, unsigned int w, unsigned int h, int channels) { return writeEXR(path, pixels, w, h, channels) ? 0 : -1; }
int vth_write_hdr(const char* path, const float* pixels, unsigned int w, unsigned int h, int stride) { return writeHDR(path, pixels, w, h, stride) ? 0 : -1; }
int vth_load_image_dims(const char* path, unsigned int* w, unsigned int* h)
{
    std::vector<float> px;
    return loadImage(path, *w, *h, px) ? 0 : -1;
}
int vth_load_image(const char* path, float* rgb)
{
    std::vector<float> px; unsigned int w, h;
    if (!loadImage(path, w, h, px)) return -1;
    memcpy(rgb, &px[0], px.size() * 4);
    return 0;
}

} // extern "C"
