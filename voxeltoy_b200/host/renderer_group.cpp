// renderer_group.cpp -- RendererGroup (vt_host.h): N Renderers, one per GPU, combined through the C ABI's vt_group.
// The reference has no counterpart (renderer/renderer.cpp:556-645 renders on one GL context); the replicated calls are the
// ones ui/glwidget.cpp makes on its Renderer, and edits follow renderer/actions.cpp:5-52 on every replica.
#include "vt_host.h"

RendererGroup::RendererGroup() : m_group(NULL) {}

RendererGroup::~RendererGroup()
{
    if (m_group) vt_group_destroy(m_group);
    for (size_t i = 0; i < m_renderers.size(); ++i) delete m_renderers[i];
}

bool RendererGroup::initialize(const std::vector<int>& cudaDevices, Mode mode)
{
    if (m_group || cudaDevices.empty()) { m_status = "RendererGroup::initialize: already initialised or no devices"; return false; }
    std::vector<vt_ctx*> ctxs;
    for (size_t i = 0; i < cudaDevices.size(); ++i) {
        Renderer* r = new Renderer();
        r->initializeOnDevice(cudaDevices[i]);
        m_renderers.push_back(r);
        if (!r->context()) { m_status = "no CUDA context on device " + std::to_string(cudaDevices[i]) + ": " + r->getStatus(); return false; }
        ctxs.push_back(r->context());
    }
    const int rc = vt_group_adopt((int)ctxs.size(), &ctxs[0], (int)mode, &m_group);
    if (rc != VT_OK) { m_status = "vt_group_adopt failed with status " + std::to_string(rc); m_group = NULL; return false; }
    return true;
}

void RendererGroup::resizeFrame(int width, int height) { forEach([&](Renderer& r) { r.resizeFrame(width, height, 0, 0, width, height); }); }
void RendererGroup::loadVoxFile(const std::string& file) { forEach([&](Renderer& r) { r.loadVoxFile(file); }); }
void RendererGroup::loadMeshAtResolution(const std::string& file, int resolution) { forEach([&](Renderer& r) { r.loadMeshAtResolution(file, resolution); }); }
void RendererGroup::setVoxelData(const vtm::V3i& resolution, const std::vector<int32_t>& voxelMaterials, const std::vector<float>& materialData,
                                 const std::vector<int32_t>& emissiveVoxelIndices)
{
    forEach([&](Renderer& r) { r.setVoxelData(resolution, voxelMaterials, materialData, emissiveVoxelIndices); });
}
void RendererGroup::updateRenderSettings(const RenderSettings& settings)
{
    forEach([&](Renderer& r) { r.renderSettings() = settings; r.updateRenderSettings(); });
}
void RendererGroup::requestAction(float x, float y, float dx, float dy, Action::PICKING_ACTION action, bool restartAccumulation)
{
    forEach([&](Renderer& r) { r.requestAction(x, y, dx, dy, action, restartAccumulation); });
}
void RendererGroup::resetRender() { forEach([](Renderer& r) { r.resetRender(); }); }
void RendererGroup::renderPasses(int nPasses) { forEach([&](Renderer& r) { r.renderPasses(nPasses); }); }

bool RendererGroup::beginCombine()
{
    if (!m_group) return false;
    if (vt_group_begin_combine(m_group) != VT_OK) { m_status = vt_group_last_error(m_group); return false; }
    return true;
}
bool RendererGroup::endCombine(float* rgbaOut)
{
    if (!m_group) return false;
    if (vt_group_end_combine(m_group, rgbaOut) != VT_OK) { m_status = vt_group_last_error(m_group); return false; }
    return true;
}
