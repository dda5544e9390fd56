// renderer.cpp -- Renderer, its services, GPUVoxelizer and the tools.
// Control flow follows renderer/renderer.cpp, renderer/actions.cpp, renderer/import.cpp,
// renderer/services/*.cpp, voxelize/gpuVoxelizer.cpp and tools/*.cpp of the reference; every place the
// reference talks to OpenGL is a call into the C ABI (include/voxeltoy_b200.h) here.
#include "vt_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>

using namespace vtm;

// =============================================================================================================
// services (renderer/services/*.cpp): the 1-vertex "compute" draws become single kernel launches
// =============================================================================================================
void RendererServiceSelectActiveVoxel::execute() { vt_pick(m_ctx, m_point.x, m_point.y); }            // servicePicking.cpp:118-127
void RendererServiceSetFocalDistance::execute() { vt_pick_focal(m_ctx, m_point.x, m_point.y); }
void RendererServiceAddVoxel::execute() { vt_add_voxel(m_ctx, m_velocity.x, -m_velocity.y); }         // serviceAddVoxel.cpp:68-79 (y flip)
void RendererServiceRemoveVoxel::execute() { vt_remove_voxel(m_ctx); }                                // serviceRemoveVoxel.cpp:46-53

// =============================================================================================================
// GPUVoxelizer (voxelize/gpuVoxelizer.cpp) + mesh normalisation (renderer/import.cpp:46-64)
// =============================================================================================================
GPUVoxelizer::GPUVoxelizer(const std::string&, Logger* logger) : m_initialized(true), m_logger(logger), m_lastMs(0) {}
GPUVoxelizer::~GPUVoxelizer() {}

bool GPUVoxelizer::voxelizeMesh(const Mesh* mesh, const M44f& meshTransform, const V3i& resolution, vt_ctx* target, int32_t fillOffset)
{
    if (!m_initialized || !mesh || !target) return false;
    vt_set_voxelize_thickness(target, (int)m_thickness);
    const int rc = vt_voxelize(target, mesh->vertices().empty() ? NULL : &mesh->vertices()[0], mesh->vertices().size() / 3,
                               mesh->indices().empty() ? NULL : &mesh->indices()[0], mesh->indices().size(),
                               &meshTransform.x[0][0], resolution.x, resolution.y, resolution.z, fillOffset);
    if (rc != VT_OK) { if (m_logger) (*m_logger)(std::string("Voxelize failed: ") + vt_last_error(target)); return false; }
    vt_get_last_voxelize_ms(target, &m_lastMs);
    return true;
}

M44f computeMeshTransform(const Box3f& bounds, const V3i& voxelResolution)
{
    // uniform scale so that the major axis spans the unit cube minus a one-voxel margin on each side
    const V3f margin(1.0f / voxelResolution.x, 1.0f / voxelResolution.y, 1.0f / voxelResolution.z);
    const int major = bounds.majorAxis();
    const float s = (float)((1.0f - 2.0 * margin[major]) / bounds.size()[major]);
    const V3f t = -bounds.min + margin / s;
    M44f m;
    m.x[0][0] = s; m.x[1][1] = s; m.x[2][2] = s;
    m.x[0][3] = t.x * s; m.x[1][3] = t.y * s; m.x[2][3] = t.z * s;
    return m;
}

// =============================================================================================================
// Renderer
// =============================================================================================================
Renderer::Renderer()                                                             // renderer.cpp:41-65
    : m_initialized(false), m_ctx(NULL), m_device(0), m_volumeResolution(16), m_numberSamples(0),
      m_currentIntegrator(INTEGRATOR_PATHTRACER), m_currentBackgroundRadianceIntegral(0), m_logger(NULL)
{
    m_camera.controller().lookAt(V3f(0, 0, 0));
    m_camera.controller().setDistanceFromTarget(100);
    m_camera.setFStop(16);
    m_renderSettings.m_imageResolution = V2i(512, 512);
    m_renderSettings.m_pathtracerMaxNumBounces = 1;
    m_renderSettings.m_pathtracerMaxSamples = 2048 * 2048;
    m_renderSettings.m_viewport[0] = m_renderSettings.m_viewport[1] = 0;
    m_renderSettings.m_viewport[2] = m_renderSettings.m_viewport[3] = 512;
    m_renderSettings.m_wireframeOpacity = 0;                                     // pathTracer.fs:40-41
    m_renderSettings.m_wireframeThickness = 0.01f;
    m_renderSettings.m_backgroundColor[0] = V3f(153.0f / 255, 187.0f / 255, 201.0f / 255) * 2.0f;   // pathTracer.fs:25-26
    m_renderSettings.m_backgroundColor[1] = V3f(77.0f / 255, 64.0f / 255, 50.0f / 255);
    m_renderSettings.m_backgroundRotationDegrees = 0;
    memset(m_services, 0, sizeof m_services);
}

Renderer::~Renderer()
{
    for (int i = 0; i < SERVICE_TOTAL; ++i) delete m_services[i];
    if (m_ctx) vt_destroy(m_ctx);
}

void Renderer::setLogger(Logger* logger) { m_logger = logger; }
void Renderer::log(const std::string& msg) { if (m_logger) (*m_logger)(msg); }
void Renderer::logTrampoline(const char* msg, void* user) { static_cast<Renderer*>(user)->log(msg); }

void Renderer::initializeOnDevice(int cudaDevice) { m_device = cudaDevice; initialize(""); }

void Renderer::initialize(const std::string& shaderPath)                          // renderer.cpp:76-147
{
    if (m_initialized) return;
    m_shaderPath = shaderPath;
    const int rc = vt_create(m_device, &m_ctx);
    if (rc != VT_OK) {
        // the reference reports a failed glewInit / shader build through the status string and stays uninitialised
        m_status = (rc == VT_ERR_NO_DEVICE) ? "no CUDA device: voxeltoy_b200 has no CPU or OpenGL fallback" : "CUDA context creation failed";
        log(m_status);
        m_ctx = NULL;
        return;
    }
    vt_set_logger(m_ctx, &Renderer::logTrampoline, this);
    m_initialized = true;
    createVoxelDataTexture(V3i(16));                                              // :98
    m_services[SERVICE_ADD_VOXEL] = new RendererServiceAddVoxel(m_ctx);           // :122-125
    m_services[SERVICE_REMOVE_VOXEL] = new RendererServiceRemoveVoxel(m_ctx);
    m_services[SERVICE_SELECT_ACTIVE_VOXEL] = new RendererServiceSelectActiveVoxel(m_ctx);
    m_services[SERVICE_SET_FOCAL_DISTANCE] = new RendererServiceSetFocalDistance(m_ctx);
    updateRenderSettings();
    updateCamera();
    m_status = "initialized";
}

void Renderer::reloadShaders(const std::string& shaderPath)                       // renderer.cpp:314-341: nothing to compile at run time
{
    m_shaderPath = shaderPath;
    if (!m_initialized) return;
    updateCamera();
    updateRenderSettings();
}

void Renderer::setIntegrator(Integrator i) { m_currentIntegrator = i; m_numberSamples = 0; if (m_initialized) updateRenderSettings(); }
void Renderer::setPartition(int mode, int rank, int world) { if (m_initialized) vt_set_partition(m_ctx, mode, rank, world); }

void Renderer::updateCamera()                                                     // renderer.cpp:343-446
{
    if (!m_initialized) return;
    const CameraParameters& cp = m_camera.parameters();
    const V3f eye = cp.eye();
    V3f right, up, forward;
    cp.getBasis(forward, right, up);

    M44f mvm, pm;                                                                 // look-at, row-major, column-vector convention
    for (int c = 0; c < 3; ++c) { mvm.x[0][c] = right[c]; mvm.x[1][c] = up[c]; mvm.x[2][c] = -forward[c]; }
    mvm.x[0][3] = -eye.dot(right); mvm.x[1][3] = -eye.dot(up); mvm.x[2][3] = eye.dot(forward);

    const float a = (float)m_renderSettings.m_imageResolution.x / m_renderSettings.m_imageResolution.y;
    memset(pm.x, 0, sizeof pm.x);
    if (cp.lensModel() == CameraParameters::CLM_ORTHOGRAPHIC) {                   // :366-385
        const float left = -std::tan(cp.fovY() / 2) * cp.distanceToTarget(), rgt = -left;
        const float bottom = -std::tan(cp.fovY() / 2) * cp.distanceToTarget() / a, top = -bottom;
        const float nearZ = -cp.nearDistance(), farZ = -cp.farDistance();
        pm.x[0][0] = 2.0f / (rgt - left); pm.x[0][3] = -(rgt + left) / (rgt - left);
        pm.x[1][1] = 2.0f / (top - bottom); pm.x[1][3] = -(top + bottom) / (top - bottom);
        pm.x[2][2] = -2.0f / (farZ - nearZ); pm.x[2][3] = -(farZ + nearZ) / (farZ - nearZ);
        pm.x[3][3] = 1;
    } else {                                                                      // :386-397
        const float n = cp.nearDistance(), f = cp.farDistance();
        const float e = 1.0f / std::tan(cp.fovY() / 2);
        pm.x[0][0] = e / a; pm.x[1][1] = e;
        pm.x[2][2] = (f + n) / (n - f); pm.x[2][3] = 2.0f * f * n / (n - f);
        pm.x[3][2] = -1;
    }
    m_mvm = mvm; m_pm = pm; m_invMvm = mvm.inverse(); m_invPm = pm.inverse();
    for (int i = 0; i < SERVICE_TOTAL; ++i)
        if (m_services[i]) m_services[i]->cameraUpdated(m_mvm, m_invMvm, m_pm, m_invPm, m_camera);

    vt_camera cam;
    memcpy(cam.inv_modelview, m_invMvm.x, sizeof cam.inv_modelview);
    memcpy(cam.proj, m_pm.x, sizeof cam.proj);
    memcpy(cam.inv_proj, m_invPm.x, sizeof cam.inv_proj);
    cam.near_z = cp.nearDistance(); cam.far_z = cp.farDistance();
    cam.lens_radius = cp.lensRadius(); cam.lens_model = (int32_t)cp.lensModel();
    vt_set_camera(m_ctx, &cam);
}

void Renderer::cameraMatrices(float invModelView[16], float proj[16], float invProj[16])
{
    memcpy(invModelView, m_invMvm.x, 64); memcpy(proj, m_pm.x, 64); memcpy(invProj, m_invPm.x, 64);
}

void Renderer::resizeFrame(int frameBufferWidth, int frameBufferHeight, int viewportX, int viewportY, int viewportW, int viewportH)
{                                                                                 // renderer.cpp:448-554
    m_numberSamples = 0;
    m_renderSettings.m_imageResolution = V2i(frameBufferWidth, frameBufferHeight);
    m_renderSettings.m_viewport[0] = viewportX; m_renderSettings.m_viewport[1] = viewportY;
    m_renderSettings.m_viewport[2] = viewportW; m_renderSettings.m_viewport[3] = viewportH;
    const float aspect = (float)frameBufferWidth / frameBufferHeight;
    if (aspect >= 1.0f) m_camera.setFilmSize(CameraParameters::FILM_SIZE_35MM, CameraParameters::FILM_SIZE_35MM / aspect);
    else m_camera.setFilmSize(CameraParameters::FILM_SIZE_35MM * aspect, CameraParameters::FILM_SIZE_35MM);
    if (!m_initialized) return;
    updateRenderSettings();      // (re)allocates the float4 accumulator: always RGBA32F (contract U4)
    updateCamera();
    for (int i = 0; i < SERVICE_TOTAL; ++i) if (m_services[i]) m_services[i]->frameResized(m_renderSettings.m_viewport);
}

Renderer::RenderResult Renderer::render() { return renderPasses(1); }             // renderer.cpp:556-645

Renderer::RenderResult Renderer::renderPasses(int nPasses)
{
    if (!m_initialized) return RR_FINISHED_RENDERING;
    processPendingActions();                                                      // :569
    const float aspect = (float)m_renderSettings.m_imageResolution.x / m_renderSettings.m_imageResolution.y;   // :571-580
    if (aspect >= 1.0f) m_camera.setFilmSize(CameraParameters::FILM_SIZE_35MM, CameraParameters::FILM_SIZE_35MM / aspect);
    else m_camera.setFilmSize(CameraParameters::FILM_SIZE_35MM * aspect, CameraParameters::FILM_SIZE_35MM);
    updateCamera();
    if (m_numberSamples == 0) vt_reset_accumulation(m_ctx);                       // average*0 in accumulation.fs: any old content is dropped
    RenderResult result = RR_SAMPLES_PENDING;
    for (int done = 0; done < nPasses;) {
        // sampleCount = min(n, maxSamples - 1) (:594): beyond maxSamples the same sample index repeats, one pass per launch
        const int first = std::min(m_numberSamples, m_renderSettings.m_pathtracerMaxSamples - 1);
        int batch = nPasses - done;
        if (m_numberSamples >= m_renderSettings.m_pathtracerMaxSamples - 1) batch = 1;
        else batch = std::min(batch, m_renderSettings.m_pathtracerMaxSamples - 1 - m_numberSamples);
        if (vt_render(m_ctx, first, batch) != VT_OK) { m_status = vt_last_error(m_ctx); return RR_FINISHED_RENDERING; }
        done += batch;
        for (int k = 0; k < batch; ++k)
            result = (m_numberSamples++ < m_renderSettings.m_pathtracerMaxSamples) ? RR_SAMPLES_PENDING : RR_FINISHED_RENDERING;   // :639-644
    }
    return result;
}

bool Renderer::onMouseMove(int dx, int dy, int buttons)                           // renderer.cpp:658-669
{
    const float ndx = (float)dx / m_renderSettings.m_imageResolution.x;
    const float ndy = (float)dy / m_renderSettings.m_imageResolution.y;
    if (m_camera.controller().onMouseMove(ndx, ndy, buttons)) { updateCamera(); m_numberSamples = 0; return true; }
    return false;
}

bool Renderer::onKeyPress(int key)                                                // renderer.cpp:670-692
{
    if (m_camera.controller().onKeyPress(key)) { updateCamera(); m_numberSamples = 0; return true; }
    if (key == vtinput::Key_Space) {
        setIntegrator((Integrator)(!(int)m_currentIntegrator));
        return true;
    }
    if (key == vtinput::Key_F) {
        m_camera.controller().focusOnBounds(m_volumeBounds);
        m_numberSamples = 0;
        updateCamera();
        return true;
    }
    return false;
}

void Renderer::createVoxelDataTexture(const V3i& resolution, const int32_t* voxelMaterials, const float* materialData,
                                      size_t materialDataSize, const int32_t* emissiveVoxelIndices, size_t numEmissiveVoxels)
{                                                                                 // renderer.cpp:833-940
    if (!m_initialized) return;
    m_volumeResolution = resolution;
    if (vt_volume_upload(m_ctx, voxelMaterials, resolution.x, resolution.y, resolution.z) != VT_OK) { m_status = vt_last_error(m_ctx); return; }
    vt_materials_upload(m_ctx, materialData, materialData ? materialDataSize : 0);
    vt_emissive_upload(m_ctx, emissiveVoxelIndices, emissiveVoxelIndices ? numEmissiveVoxels : 0);
    m_materialData.assign(materialData, materialData + (materialData ? materialDataSize : 0));
    float bmin[3], bmax[3];
    vt_get_volume_info(m_ctx, NULL, bmin, bmax, NULL);                            // longest side 1000, centred (:845-850)
    m_volumeBounds = Box3f(V3f(bmin[0], bmin[1], bmin[2]), V3f(bmax[0], bmax[1], bmax[2]));
    for (int i = 0; i < SERVICE_TOTAL; ++i) if (m_services[i]) m_services[i]->volumeReloaded(m_volumeResolution, m_volumeBounds);
}

void Renderer::resetRender() { updateCamera(); m_numberSamples = 0; }             // renderer.cpp:941-945

bool Renderer::loadBackgroundImage(float& mapIntegralTimesSin)                    // renderer.cpp:947-1055
{
    if (m_renderSettings.m_backgroundImage.empty()) { mapIntegralTimesSin = 0; return true; }
    if (m_renderSettings.m_backgroundImage == m_currentBackgroundImage) { mapIntegralTimesSin = m_currentBackgroundRadianceIntegral; return true; }
    unsigned int w, h;
    std::vector<float> pixels;
    if (!loadImage(m_renderSettings.m_backgroundImage, w, h, pixels)) {
        log("Background image loading failed: " + m_renderSettings.m_backgroundImage);
        return false;
    }
    // renderer.cpp:1004-1046 + image.cpp:68-389 (importance function, CDFs, integral) run on the device: vt_env_build
    float integral = 0;
    if (vt_env_build(m_ctx, &pixels[0], (int)w, (int)h) != VT_OK || vt_get_env_info(m_ctx, NULL, &integral, NULL) != VT_OK) {
        m_status = vt_last_error(m_ctx); log("Background CDF construction failed: " + m_status); return false;
    }
    mapIntegralTimesSin = integral;
    m_currentBackgroundImage = m_renderSettings.m_backgroundImage;
    m_currentBackgroundRadianceIntegral = integral;
    return true;
}

void Renderer::updateRenderSettings()                                             // renderer.cpp:1057-1106
{
    if (!m_initialized) return;
    float integral = 0;
    if (!loadBackgroundImage(integral)) return;
    vt_settings st;
    st.width = m_renderSettings.m_imageResolution.x; st.height = m_renderSettings.m_imageResolution.y;
    st.max_bounces = m_renderSettings.m_pathtracerMaxNumBounces;
    st.integrator = (int)m_currentIntegrator;
    for (int i = 0; i < 3; ++i) { st.bg_top[i] = m_renderSettings.m_backgroundColor[0][i]; st.bg_bottom[i] = m_renderSettings.m_backgroundColor[1][i]; }
    st.use_env_image = m_renderSettings.m_backgroundImage.empty() ? 0 : 1;
    st.env_rotation_rad = (float)(m_renderSettings.m_backgroundRotationDegrees * M_PI / 180.0f);
    st.wireframe_opacity = m_renderSettings.m_wireframeOpacity;
    st.wireframe_thickness = m_renderSettings.m_wireframeThickness;
    if (vt_set_settings(m_ctx, &st) != VT_OK) { m_status = vt_last_error(m_ctx); return; }
    resetRender();
}

bool Renderer::readAverage(float* rgbaOut)
{
    if (!m_initialized) return false;
    return vt_read_average(m_ctx, rgbaOut) == VT_OK;
}

void Renderer::saveImage(const std::string& file)                                 // renderer.cpp:1108-1140
{
    if (!m_initialized) return;
    const int xres = m_renderSettings.m_imageResolution.x, yres = m_renderSettings.m_imageResolution.y;
    const std::string ext = file.size() > 4 ? file.substr(file.size() - 4) : std::string();
    if (ext == ".png" || ext == ".ppm") {
        // 8-bit formats: conversion + vertical flip (the reference's negative stride, :1131-1136) on the device, 4 B/pixel read back
        std::vector<unsigned char> px((size_t)xres * yres * 4);
        if (vt_read_display(m_ctx, &px[0], 1) != VT_OK) { m_status = vt_last_error(m_ctx); return; }
        if (ext == ".png") { if (!writePNG(file, &px[0], xres, yres)) log("could not write " + file); return; }
        FILE* fp = fopen(file.c_str(), "wb");
        if (!fp) return;
        fprintf(fp, "P6\n%d %d\n255\n", xres, yres);
        for (size_t i = 0; i < (size_t)xres * yres; ++i) fwrite(&px[4 * i], 1, 3, fp);
        fclose(fp);
        return;
    }
    std::vector<float> pixels((size_t)xres * yres * 4);
    if (!readAverage(&pixels[0])) return;
    if (ext == ".exr" || ext == ".hdr") {                                      // float formats, rows top-down: the flip of :1131-1136 on the host
        std::vector<float> top((size_t)xres * yres * 4);
        for (int y = 0; y < yres; ++y) memcpy(&top[(size_t)y * xres * 4], &pixels[(size_t)(yres - 1 - y) * xres * 4], (size_t)xres * 4 * sizeof(float));
        const bool ok = ext == ".exr" ? writeEXR(file, &top[0], xres, yres, 4) : writeHDR(file, &top[0], xres, yres, 4);
        if (!ok) log("could not write " + file);
        return;
    }
    FILE* fp = fopen(file.c_str(), "wb");
    if (!fp) return;
    fprintf(fp, "PF\n%d %d\n-1.0\n", xres, yres);                              // PFM stores bottom-to-top = GL order
    for (int y = 0; y < yres; ++y)
        for (int x = 0; x < xres; ++x) fwrite(&pixels[((size_t)y * xres + x) * 4], sizeof(float), 3, fp);
    fclose(fp);
}

std::vector<Material::SerializedData> Renderer::getMaterials() const             // renderer.cpp:1142-1203
{
    std::vector<Material::SerializedData> result;
    if (!m_initialized) return result;
    std::vector<float> storage(m_materialData.size());
    if (!storage.empty()) vt_read_materials(m_ctx, &storage[0], storage.size());  // read back from the device, like glGetTexImage
    size_t offset = 0;
    while (offset < storage.size()) {
        const int mt = (int)storage[offset];
        size_t dataSize = 0;
        if (mt == Material::MT_LAMBERT && offset + 6 < storage.size() + 0) {
            Material::LambertMaterialData d; memcpy(&d, &storage[offset + 1], sizeof d);
            result.push_back(Material::serializeLambert(d, offset)); dataSize = 6;
        } else if (mt == Material::MT_METAL && offset + 7 < storage.size()) {
            Material::MetalMaterialData d; memcpy(&d, &storage[offset + 1], sizeof d);
            result.push_back(Material::serializeMetal(d, offset)); dataSize = 7;
        } else if (mt == Material::MT_PLASTIC && offset + 7 < storage.size()) {
            Material::PlasticMaterialData d; memcpy(&d, &storage[offset + 1], sizeof d);
            result.push_back(Material::serializePlastic(d, offset)); dataSize = 7;
        } else {
            if (m_logger) (*m_logger)("Unrecognized material type. Aborting.");
            break;
        }
        offset += 1 + dataSize;
    }
    return result;
}

void Renderer::updateMaterialColor(unsigned int dataOffset, const float color[3])  // renderer.cpp:1205-1219
{
    if (!m_initialized) return;
    if (vt_material_update(m_ctx, dataOffset, color, 3) == VT_OK && dataOffset + 3 <= m_materialData.size())
        memcpy(&m_materialData[dataOffset], color, 3 * sizeof(float));
}
void Renderer::updateMaterialValue(unsigned int dataOffset, float value)           // renderer.cpp:1221-1235
{
    if (!m_initialized) return;
    if (vt_material_update(m_ctx, dataOffset, &value, 1) == VT_OK && dataOffset < m_materialData.size()) m_materialData[dataOffset] = value;
}

// ---- renderer/actions.cpp ---------------------------------------------------------------------------------------
void Renderer::requestAction(float x, float y, float dx, float dy, Action::PICKING_ACTION action, bool restartAccumulation)
{
    Action a;
    a.m_point = V2f(x, y); a.m_velocity = V2f(dx, dy); a.m_type = action; a.m_invalidatesRender = restartAccumulation;
    m_scheduledActions.push_back(a);
}

void Renderer::processPendingActions()                                            // actions.cpp:20-52
{
    for (size_t i = 0; i < m_scheduledActions.size(); ++i) {
        const Action& a = m_scheduledActions[i];
        if (a.m_invalidatesRender) m_numberSamples = 0;
        V2f position(a.m_point.x * m_renderSettings.m_imageResolution.x, (1.0f - a.m_point.y) * m_renderSettings.m_imageResolution.y);
        V2f velocity(a.m_velocity.x * m_renderSettings.m_imageResolution.x, a.m_velocity.y * m_renderSettings.m_imageResolution.y);
        RendererServiceType s = SERVICE_TOTAL;
        switch (a.m_type) {
        case Action::PA_SELECT_FOCAL_POINT: s = SERVICE_SET_FOCAL_DISTANCE; break;
        case Action::PA_SELECT_ACTIVE_VOXEL: s = SERVICE_SELECT_ACTIVE_VOXEL; break;
        case Action::PA_ADD_VOXEL: s = SERVICE_ADD_VOXEL; break;
        case Action::PA_REMOVE_VOXEL: s = SERVICE_REMOVE_VOXEL; break;
        default: break;
        }
        if (s != SERVICE_TOTAL && m_services[s]) { m_services[s]->setMouseParameters(position, velocity); m_services[s]->execute(); }
    }
    m_scheduledActions.clear();
}

// ---- renderer/import.cpp ------------------------------------------------------------------------------------------
void Renderer::setVoxelData(const V3i& resolution, const std::vector<int32_t>& voxelMaterials, const std::vector<float>& materialData,
                            const std::vector<int32_t>& emissiveVoxelIndices)
{
    if (!m_initialized) return;
    if (voxelMaterials.size() != (size_t)resolution.x * resolution.y * resolution.z) { log("setVoxelData: grid size does not match the resolution"); return; }
    std::vector<int32_t> emissive(emissiveVoxelIndices);
    V3i res = resolution;
    pruneInteriorEmissiveVoxels(voxelMaterials, res, emissive);
    createVoxelDataTexture(res, &voxelMaterials[0], materialData.empty() ? NULL : &materialData[0], materialData.size(),
                           emissive.empty() ? NULL : &emissive[0], emissive.size());
    m_camera.controller().setDistanceFromTarget(m_volumeBounds.size().length() * 0.5f);   // import.cpp:41
    resetRender();
}

void Renderer::loadVoxFile(const std::string& file)                               // import.cpp:11-44
{
    if (!m_initialized) return;
    std::vector<int32_t> voxelMaterials, emissive;
    std::vector<float> materialData;
    MagicaVoxelLoader loader;
    loader.setPaletteRules(m_voxPaletteRules);
    V3i res;
    if (!loader.load(file, voxelMaterials, materialData, emissive, res)) { log("[error] MV_VoxelModel :: " + loader.m_error); return; }
    setVoxelData(res, voxelMaterials, materialData, emissive);
}

void Renderer::loadMesh(const std::string& file) { loadMeshAtResolution(file, 64); }   // import.cpp:66-131 (hard-coded 64^3 :75)

void Renderer::loadMeshAtResolution(const std::string& file, int resolution)
{
    if (!m_initialized) return;
    Mesh* mesh = MeshLoader::loadFromOBJ(file.c_str());
    if (mesh == NULL) return;
    const V3i res(resolution);
    // one default Lambert material (grey) at offset 0 for the voxelized shell (contract U5: the reference leaves the
    // material texture empty and writes offset 1 into an uninitialised grid)
    std::vector<float> materialData;
    VoxLoader::generateMaterialLambert(V3f(0, 0, 0), V3f(0.5f, 0.5f, 0.5f), materialData);
    const M44f meshTransform = computeMeshTransform(mesh->bounds(), res);
    GPUVoxelizer voxelizer(m_shaderPath, m_logger);
    if (voxelizer.voxelizeMesh(mesh, meshTransform, res, m_ctx, 0)) {
        m_volumeResolution = res;
        vt_materials_upload(m_ctx, &materialData[0], materialData.size());
        vt_emissive_upload(m_ctx, NULL, 0);
        m_materialData = materialData;
        float bmin[3], bmax[3];
        vt_get_volume_info(m_ctx, NULL, bmin, bmax, NULL);
        m_volumeBounds = Box3f(V3f(bmin[0], bmin[1], bmin[2]), V3f(bmax[0], bmax[1], bmax[2]));
        for (int i = 0; i < SERVICE_TOTAL; ++i) if (m_services[i]) m_services[i]->volumeReloaded(m_volumeResolution, m_volumeBounds);
    }
    delete mesh;
    resetRender();
}

void Renderer::pruneInteriorEmissiveVoxels(const std::vector<int32_t>& voxelMaterials, V3i& res, std::vector<int32_t>& emissive)
{                                                                                 // import.cpp:133-203
    if (emissive.empty()) return;
    const size_t numInput = emissive.size();
    static const int nb[6][3] = { {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1} };
    size_t pruned = 0;
    for (int i = 0; i < (int)emissive.size(); ++i) {
        const int32_t voxel = emissive[i];
        const int z = voxel / (res.x * res.y);
        const int rem = voxel - z * res.x * res.y;
        const int y = rem / res.x, x = rem - y * res.x;
        bool visible = false;
        for (int k = 0; k < 6 && !visible; ++k) {
            const int nx = x + nb[k][0], ny = y + nb[k][1], nz = z + nb[k][2];
            const bool outside = nx < 0 || nx >= res.x || ny < 0 || ny >= res.y || nz < 0 || nz >= res.z;
            if (outside || voxelMaterials[(size_t)nx + (size_t)ny * res.x + (size_t)nz * res.x * res.y] < 0) visible = true;
        }
        if (!visible) {                                                           // swap-with-last removal keeps the reference's list order
            emissive[i] = emissive.back();
            emissive.pop_back();
            --i; ++pruned;
        }
    }
    std::ostringstream ss;
    ss << "Pruned emissive voxels: " << pruned << "/" << numInput << " (" << (float)pruned / numInput * 100 << "%)";
    log(ss.str());
}

// =============================================================================================================
// tools (tools/toolAddRemoveVoxel.cpp, tools/toolFocalDistance.cpp)
// =============================================================================================================
bool ToolAddRemoveVoxel::mousePressEvent(const MouseEvent* event, WidgetSize wd)
{
    bool res = false;
    const int dx = event->x - m_lastPos[0], dy = event->y - m_lastPos[1];
    const float fx = (float)event->x / wd.width, fy = (float)event->y / wd.height;
    const float fdx = (float)dx / wd.width, fdy = (float)dy / wd.height;
    if (event->buttons & vtinput::LeftButton) {
        m_renderer.requestAction(fx, fy, fdx, fdy, (event->modifiers & vtinput::ControlModifier) ? Action::PA_REMOVE_VOXEL : Action::PA_ADD_VOXEL, true);
        res = true;
    }
    m_lastPos[0] = event->x; m_lastPos[1] = event->y;
    return res;
}
bool ToolAddRemoveVoxel::mouseMoveEvent(const MouseEvent* event, WidgetSize wd)
{
    const int dx = event->x - m_lastPos[0], dy = event->y - m_lastPos[1];
    const float fx = (float)event->x / wd.width, fy = (float)event->y / wd.height;
    const float fdx = (float)dx / wd.width, fdy = (float)dy / wd.height;
    m_renderer.requestAction(fx, fy, fdx, fdy, Action::PA_SELECT_ACTIVE_VOXEL, true);
    if (event->buttons & vtinput::LeftButton)
        m_renderer.requestAction(fx, fy, fdx, fdy, (event->modifiers & vtinput::ControlModifier) ? Action::PA_REMOVE_VOXEL : Action::PA_ADD_VOXEL, true);
    m_lastPos[0] = event->x; m_lastPos[1] = event->y;
    return false;                         // the camera controller still sees the event
}
bool ToolFocalDistance::mousePressEvent(const MouseEvent* event, WidgetSize wd)
{
    if (!(event->buttons & vtinput::LeftButton)) return false;
    m_renderer.requestAction((float)event->x / wd.width, (float)event->y / wd.height, 0, 0, Action::PA_SELECT_FOCAL_POINT, true);
    return true;
}
