// vt_host.h -- host-side classes of voxeltoy_b200.
//
// These keep the class-level API of the reference's C++ host code (class and method names, argument
// meaning, "no-op before initialize" / "return false + log" error behaviour) so that code written
// against voxelToy's Renderer / GPUVoxelizer / Camera / loaders / services / tools compiles against
// this header instead. Everything that was an OpenGL call in the reference is a call into the C ABI
// (include/voxeltoy_b200.h); there is no GL, Qt, Imath, OpenImageIO or boost dependency and no CPU
// rendering fallback. Reference locations are cited per class (paths relative to /root/reference/src).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/voxeltoy_b200.h"

// ---- minimal vector types standing in for Imath (only what the class API exposes) -------------------
namespace vtm {
struct V2f { float x, y; V2f() : x(0), y(0) {} V2f(float a, float b) : x(a), y(b) {} explicit V2f(float a) : x(a), y(a) {} };
struct V2i { int x, y; V2i() : x(0), y(0) {} V2i(int a, int b) : x(a), y(b) {} };
struct V3i { int x, y, z; V3i() : x(0), y(0), z(0) {} V3i(int a, int b, int c) : x(a), y(b), z(c) {} explicit V3i(int a) : x(a), y(a), z(a) {} };
struct V3f {
    float x, y, z;
    V3f() : x(0), y(0), z(0) {}
    V3f(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3f(float a) : x(a), y(a), z(a) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    V3f operator+(const V3f& o) const { return V3f(x + o.x, y + o.y, z + o.z); }
    V3f operator-(const V3f& o) const { return V3f(x - o.x, y - o.y, z - o.z); }
    V3f operator-() const { return V3f(-x, -y, -z); }
    V3f operator*(float s) const { return V3f(x * s, y * s, z * s); }
    V3f operator/(float s) const { return V3f(x / s, y / s, z / s); }
    float dot(const V3f& o) const { return x * o.x + y * o.y + z * o.z; }
    V3f cross(const V3f& o) const { return V3f(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    float length() const;
    V3f normalized() const;
};
inline V3f operator*(float s, const V3f& v) { return v * s; }
struct Box3f {
    V3f min, max;
    Box3f() { makeEmpty(); }
    Box3f(const V3f& a, const V3f& b) : min(a), max(b) {}
    void makeEmpty();
    void extendBy(const V3f& p);
    V3f size() const { return max - min; }
    V3f center() const { return (max + min) * 0.5f; }
    int majorAxis() const;
};
struct M44f {
    float x[4][4];
    M44f() { makeIdentity(); }
    void makeIdentity();
    M44f inverse() const;                 // general Gauss-Jordan in double, rounded to float
    M44f operator*(const M44f& o) const;
};
} // namespace vtm

// ---- log/logger.h:7-11 ------------------------------------------------------------------------------
class Logger {
public:
    virtual ~Logger() {}
    virtual void operator()(const std::string& msg) = 0;
};

// ---- renderer/material/material.h:8-51 ----------------------------------------------------------------
namespace Material {
enum MaterialType { MT_LAMBERT = 0, MT_METAL = 1, MT_PLASTIC = 2 };
struct LambertMaterialData { float emission[3]; float albedo[3]; };
struct MetalMaterialData { float emission[3]; float reflectance[3]; float roughness; };
struct PlasticMaterialData { float emission[3]; float diffuseAlbedo[3]; float roughness; };
struct SerializedData {
    std::string m_propertyName;
    float m_dataRangeFrom, m_dataRangeTo;
    float m_value;
    size_t m_dataOffset;
    enum { PROPERTY_TYPE_FLOAT, PROPERTY_TYPE_COLOR } m_propertyType;
    std::vector<SerializedData> m_childProperties;
};
SerializedData serializeLambert(const LambertMaterialData&, size_t);
SerializedData serializeMetal(const MetalMaterialData&, size_t);
SerializedData serializePlastic(const PlasticMaterialData&, size_t);
} // namespace Material

// ---- renderer/loaders/voxLoader.h:9-52, magicaVoxel.h:5-12 ----------------------------------------------
class VoxLoader {
public:
    virtual ~VoxLoader() {}
    // voxelMaterials: dense grid (x fastest) of offsets into materialData, -1 = empty.
    // emissiveVoxelIndices: linear index of every voxel whose material emits light.
    virtual bool load(const std::string& filePath, std::vector<int32_t>& voxelMaterials, std::vector<float>& materialData,
                      std::vector<int32_t>& emissiveVoxelIndices, vtm::V3i& voxelResolution) = 0;
    // public in this build (protected in the reference) so that synthetic scenes can author materials
    static void generateMaterialLambert(vtm::V3f emission, vtm::V3f albedo, std::vector<float>& materialData);
    static void generateMaterialMetal(vtm::V3f emission, vtm::V3f reflectance, float roughness, std::vector<float>& materialData);
    static void generateMaterialPlastic(vtm::V3f emission, vtm::V3f diffuseAlbedo, float roughness, std::vector<float>& materialData);
    static float getMaterialEmisiveness(const float* materialData);
};
class MagicaVoxelLoader : public VoxLoader {
public:
    // New-build extension: the .vox format carries colours only, and the reference can therefore generate
    // nothing but Lambert records (magicaVoxel.cpp:306-317). A palette rule turns one colour index into a Lambert / Metal /
    // Plastic record with emission and roughness; the colour stays the palette's. Without rules the loader is the reference's.
    struct PaletteRule { int colorIndex; Material::MaterialType type; vtm::V3f emission; float roughness; };
    void setPaletteRules(const std::vector<PaletteRule>& rules) { m_rules = rules; }
    const std::vector<PaletteRule>& paletteRules() const { return m_rules; }
    bool load(const std::string& filePath, std::vector<int32_t>& voxelMaterials, std::vector<float>& materialData,
              std::vector<int32_t>& emissiveVoxelIndices, vtm::V3i& voxelResolution) override;
    bool loadFromMemory(const unsigned char* bytes, size_t n, std::vector<int32_t>& voxelMaterials, std::vector<float>& materialData,
                        std::vector<int32_t>& emissiveVoxelIndices, vtm::V3i& voxelResolution);
    std::string m_error;
private:
    std::vector<PaletteRule> m_rules;
};

// ---- mesh/mesh.h, mesh/meshLoader.h:8-15 ------------------------------------------------------------------
class Mesh {
public:
    Mesh(const float* vertices, size_t numVertices, const unsigned int* indices, size_t numIndices);
    const vtm::Box3f& bounds() const { return m_bounds; }
    const std::vector<float>& vertices() const { return m_vertices; }
    const std::vector<unsigned int>& indices() const { return m_indices; }
private:
    std::vector<float> m_vertices;
    std::vector<unsigned int> m_indices;
    vtm::Box3f m_bounds;
};
vtm::Box3f computeBounds(const float* vertices, size_t numVertices);
class MeshLoader {
public:
    static void loadFromOBJ(const char* filePath, std::vector<float>& vertices, std::vector<unsigned int>& indices);
    static Mesh* loadFromOBJ(const char* filePath);
    static bool loadFromOBJMemory(const char* text, size_t n, std::vector<float>& vertices, std::vector<unsigned int>& indices);
};

// ---- camera/cameraParameters.h:18-129, camera.h:8-35, cameraController.h, orbit/fly ------------------------
class CameraParameters {
public:
    CameraParameters();
    enum CameraLensModel { CLM_PINHOLE, CLM_THIN_LENS, CLM_ORTHOGRAPHIC };
    const vtm::V3f& eye() const { return m_eye; }
    const vtm::V3f& target() const { return m_target; }
    float distanceToTarget() const;
    void getBasis(vtm::V3f& forwardUnitVector, vtm::V3f& rightUnitVector, vtm::V3f& upUnitVector) const;
    vtm::V3f forwardUnitVector() const;
    vtm::V3f rightUnitVector() const;
    vtm::V3f upUnitVector() const;
    float rotationTheta() const;
    float rotationPhi() const;
    float fovY() const { return m_fovY; }
    float focalLength() const;
    float nearDistance() const { return m_near; }
    float farDistance() const { return m_far; }
    float focalDistance() const { return m_focalDistance; }
    vtm::V2f filmSize() const { return m_filmSize; }
    float lensRadius() const { return m_lensRadius; }
    CameraLensModel lensModel() const { return m_lensModel; }
    static float FILM_SIZE_35MM;
private:
    friend class CameraController;
    friend class Camera;
    friend class OrbitCameraController;
    friend class FlyCameraController;
    void lookAt(const vtm::V3f& target);
    void setDistanceFromTarget(float distance);
    void orbitAroundTarget(float theta, float phi);
    void orbitAroundEye(float theta, float phi);
    void setEyeTarget(const vtm::V3f& eye, const vtm::V3f& target);
    void setFovY(float fov);
    void setNearDistance(float d);
    void setFarDistance(float d);
    void setFocalLength(float length);
    void setFocalDistance(float distance);
    void setFilmSize(float filmW, float filmH);
    void setLensRadius(float radius);
    void setFStop(float fstop);
    void setLensModel(CameraLensModel model);
    vtm::V3f m_target, m_eye;
    float m_fovY, m_near, m_far, m_focalDistance, m_lensRadius;
    vtm::V2f m_filmSize;
    CameraLensModel m_lensModel;
};

// plain replacements for the Qt enums the reference's controllers read (Qt::MouseButton, Qt::Key)
namespace vtinput {
enum MouseButton { LeftButton = 1, RightButton = 2, MiddleButton = 4 };
enum KeyModifier { ControlModifier = 0x04000000 };
enum Key { Key_Space = 0x20, Key_A = 0x41, Key_D = 0x44, Key_F = 0x46, Key_S = 0x53, Key_W = 0x57 };
}

class CameraController {
public:
    virtual ~CameraController() {}
    CameraController(CameraParameters* parameters) : m_parameters(parameters) {}
    virtual bool onMouseMove(float dx, float dy, int buttons) = 0;
    virtual bool onKeyPress(int key) = 0;
    void lookAt(const vtm::V3f& target);
    void setDistanceFromTarget(float distance);
    void focusOnBounds(const vtm::Box3f& bounds);
    // new-build convenience used by headless drivers (the UI reaches these through mouse deltas)
    void orbitAroundTarget(float theta, float phi);
protected:
    CameraParameters* m_parameters;
};
class OrbitCameraController : public CameraController {
public:
    OrbitCameraController(CameraParameters* p) : CameraController(p) {}
    bool onMouseMove(float dx, float dy, int buttons) override;
    bool onKeyPress(int key) override;
};
class FlyCameraController : public CameraController {
public:
    FlyCameraController(CameraParameters* p) : CameraController(p) {}
    bool onMouseMove(float dx, float dy, int buttons) override;
    bool onKeyPress(int key) override;
};
class Camera {
public:
    Camera();
    ~Camera();
    const CameraParameters& parameters() const { return m_parameters; }
    CameraController& controller() const { return *m_controller; }
    void setLensModel(CameraParameters::CameraLensModel);
    void setFocalLength(float length);
    void setFocalDistance(float distance);
    void setFilmSize(float filmW, float filmH);
    void setLensRadius(float radius);
    void setFStop(float fstop);
    enum CameraControllerMode { CCM_ORBIT, CCM_FLY };
    void setCameraController(CameraControllerMode);
private:
    Camera(const Camera&);
    Camera& operator=(const Camera&);
    CameraParameters m_parameters;
    CameraController* m_controller;
    CameraControllerMode m_controllerMode;
};

// ---- renderer/renderSettings.h:6-29, renderer/actions.h:5-23 ---------------------------------------------------
struct RenderSettings {
    int m_pathtracerMaxNumBounces;
    int m_pathtracerMaxSamples;
    vtm::V2i m_imageResolution;
    int m_viewport[4];
    float m_wireframeOpacity;
    float m_wireframeThickness;
    std::string m_backgroundImage;       // path of a Radiance .hdr or .pfm file ("" = gradient)
    vtm::V3f m_backgroundColor[2];       // gradient (top / bottom)
    int m_backgroundRotationDegrees;
};
struct Action {
    enum PICKING_ACTION { PA_SELECT_FOCAL_POINT, PA_SELECT_ACTIVE_VOXEL, PA_ADD_VOXEL, PA_REMOVE_VOXEL };
    PICKING_ACTION m_type;
    vtm::V2f m_point;      // normalised window coordinates, origin top-left
    vtm::V2f m_velocity;
    bool m_invalidatesRender;
};

// ---- renderer/image.h: environment map -> CDFs (image.cpp:68-283, 349-389) --------------------------------------
bool loadImage(const std::string& path, unsigned int& outWidth, unsigned int& outHeight, std::vector<float>& outPixelData);
bool calculateCDF(const float* rgbPixels, unsigned int imageWidth, unsigned int imageHeight,
                  std::vector<float>& cdfUData, unsigned int& cdfUDataWidth, unsigned int& cdfUDataHeight,
                  std::vector<float>& cdfVData, float& environmentTextureIntegral);
bool writePFM(const std::string& path, const float* rgb, unsigned int w, unsigned int h);
bool writePNG(const std::string& path, const unsigned char* rgba8, unsigned int w, unsigned int h);   // rows top-down, zlib deflate
bool writeHDR(const std::string& path, const float* pixels, unsigned int w, unsigned int h, int stride);   // Radiance RGBE, rows top-down
// image_formats.cpp: OpenEXR (scanline; NONE / RLE / ZIPS / ZIP; HALF / FLOAT / UINT) and PNG readers, OpenEXR writer (FLOAT, ZIP); rows top-down
bool readEXR(const std::string& path, unsigned int& outWidth, unsigned int& outHeight, std::vector<float>& rgb, std::string* why = nullptr);
bool readPNG(const std::string& path, unsigned int& outWidth, unsigned int& outHeight, std::vector<float>& rgb, std::string* why = nullptr);
bool writeEXR(const std::string& path, const float* pixels, unsigned int w, unsigned int h, int channels);

// ---- renderer/services/service.h:13-77 ------------------------------------------------------------------------------
class Renderer;
class RendererService {
public:
    explicit RendererService(vt_ctx* ctx) : m_ctx(ctx) {}
    virtual ~RendererService() {}
    virtual bool reload(const std::string& /*shaderPath*/, Logger* /*logger*/) { return true; }   // kernels are AOT-compiled
    virtual void cameraUpdated(const vtm::M44f&, const vtm::M44f&, const vtm::M44f&, const vtm::M44f&, const Camera&) {}
    virtual void frameResized(int /*viewport*/[4]) {}
    virtual void volumeReloaded(const vtm::V3i&, const vtm::Box3f&) {}
    virtual void setMouseParameters(vtm::V2f& point, vtm::V2f& velocity) { m_point = point; m_velocity = velocity; }
    virtual void execute() = 0;
protected:
    vt_ctx* m_ctx;
    vtm::V2f m_point, m_velocity;
};
enum RendererServiceType { SERVICE_ADD_VOXEL = 0, SERVICE_REMOVE_VOXEL, SERVICE_SELECT_ACTIVE_VOXEL, SERVICE_SET_FOCAL_DISTANCE, SERVICE_TOTAL };
class RendererServiceAddVoxel : public RendererService { public: using RendererService::RendererService; void execute() override; };
class RendererServiceRemoveVoxel : public RendererService { public: using RendererService::RendererService; void execute() override; };
class RendererServiceSelectActiveVoxel : public RendererService { public: using RendererService::RendererService; void execute() override; };
class RendererServiceSetFocalDistance : public RendererService { public: using RendererService::RendererService; void execute() override; };

// ---- voxelize/gpuVoxelizer.h:9-25 -------------------------------------------------------------------------------------
class GPUVoxelizer {
public:
    GPUVoxelizer(const std::string& shaderPath, Logger* logger = NULL);   // shaderPath is accepted and ignored
    ~GPUVoxelizer();
    // replaces the reference's `GLuint textureUnit` (the bound R32I texture) by the context that owns the grid
    bool voxelizeMesh(const Mesh* mesh, const vtm::M44f& meshTransform, const vtm::V3i& resolution, vt_ctx* target, int32_t fillOffset = 0);
    float lastMilliseconds() const { return m_lastMs; }
    // voxelize.gs:15-19: the reference selects THIN / FAT by editing `#define THICKNESS`; here it is a property of the voxelizer
    enum Thickness { THIN = VT_VOXELIZE_THIN, FAT = VT_VOXELIZE_FAT };
    void setThickness(Thickness t) { m_thickness = t; }
private:
    bool m_initialized;
    Logger* m_logger;
    float m_lastMs;
    Thickness m_thickness = THIN;
};
vtm::M44f computeMeshTransform(const vtm::Box3f& bounds, const vtm::V3i& voxelResolution);   // renderer/import.cpp:46-64

// ---- renderer/renderer.h:23-122 ------------------------------------------------------------------------------------------
class Renderer {
public:
    Renderer();
    ~Renderer();
    void initialize(const std::string& shaderPath);            // shaderPath ignored (no run-time shader compilation)
    void initializeOnDevice(int cudaDevice);                   // new-build: pick the GPU (initialize() uses device 0)
    void resizeFrame(int frameBufferWidth, int frameBufferHeight, int viewportX, int viewportY, int viewportW, int viewportH);
    enum RenderResult { RR_SAMPLES_PENDING, RR_FINISHED_RENDERING };
    RenderResult render();
    RenderResult renderPasses(int nPasses);                    // new-build: n progressive passes in one launch
    void reloadShaders(const std::string& shaderPath);
    void loadMesh(const std::string& file);
    void loadMeshAtResolution(const std::string& file, int resolution);   // new-build: the reference hard-codes 64 (import.cpp:75)
    void loadVoxFile(const std::string& file);
    void setVoxPaletteRules(const std::vector<MagicaVoxelLoader::PaletteRule>& rules) { m_voxPaletteRules = rules; }   // new-build: applied by loadVoxFile
    // new-build: the public twin of createVoxelDataTexture for scenes that are not files
    void setVoxelData(const vtm::V3i& resolution, const std::vector<int32_t>& voxelMaterials, const std::vector<float>& materialData,
                      const std::vector<int32_t>& emissiveVoxelIndices);
    void pruneInteriorEmissiveVoxels(const std::vector<int32_t>& voxelMaterials, vtm::V3i& volumeResolution, std::vector<int32_t>& emissiveVoxelIndices);
    void saveImage(const std::string& file);                   // .pfm (float RGB) or .png / .ppm (8 bit, vt_read_display), vertically flipped like the reference
    bool readAverage(float* rgbaOut);                          // GL orientation (row 0 = bottom)
    void resetRender();
    bool onMouseMove(int dx, int dy, int buttons);
    bool onKeyPress(int key);
    Camera& camera() { return m_camera; }
    RenderSettings& renderSettings() { return m_renderSettings; }
    void updateRenderSettings();
    const std::string& getStatus() const { return m_status; }
    void setLogger(Logger* logger);
    void requestAction(float x, float y, float dx, float dy, Action::PICKING_ACTION action, bool restartAccumulation);
    enum Integrator { INTEGRATOR_PATHTRACER = 0, INTEGRATOR_EDIT_MODE, INTEGRATOR_TOTAL };
    std::vector<Material::SerializedData> getMaterials() const;
    void updateMaterialColor(unsigned int dataOffset, const float color[3]);
    void updateMaterialValue(unsigned int dataOffset, float value);
    // new-build accessors
    vt_ctx* context() const { return m_ctx; }
    int numberSamples() const { return m_numberSamples; }
    const vtm::Box3f& volumeBounds() const { return m_volumeBounds; }
    const vtm::V3i& volumeResolution() const { return m_volumeResolution; }
    void setIntegrator(Integrator i);
    void setPartition(int mode, int rank, int world);
    void cameraMatrices(float invModelView[16], float proj[16], float invProj[16]);
private:
    void updateCamera();
    void createVoxelDataTexture(const vtm::V3i& resolution, const int32_t* voxelMaterials = NULL, const float* materialData = NULL,
                                size_t materialDataSize = 0, const int32_t* emissiveVoxelIndices = NULL, size_t numEmissiveVoxels = 0);
    bool loadBackgroundImage(float&);
    void processPendingActions();
    void log(const std::string& msg);
    static void logTrampoline(const char* msg, void* user);

    bool m_initialized;
    vt_ctx* m_ctx;
    int m_device;
    vtm::Box3f m_volumeBounds;
    vtm::V3i m_volumeResolution;
    int m_numberSamples;
    Camera m_camera;
    RenderSettings m_renderSettings;
    std::string m_shaderPath, m_status;
    std::vector<Action> m_scheduledActions;
    Integrator m_currentIntegrator;
    std::string m_currentBackgroundImage;
    float m_currentBackgroundRadianceIntegral;
    Logger* m_logger;
    RendererService* m_services[SERVICE_TOTAL];
    std::vector<float> m_materialData;          // host mirror for getMaterials (the reference reads the texture back)
    std::vector<MagicaVoxelLoader::PaletteRule> m_voxPaletteRules;
    vtm::M44f m_mvm, m_invMvm, m_pm, m_invPm;
};

// ---- RendererGroup: one frame on several GPUs (new-build; the reference's Renderer owns a single GL context) ----------------
// N replicas of the scene, one Renderer (and one vt_ctx) per CUDA device, behind the calls the UI makes on one Renderer:
// scene / settings / camera changes and edit actions go to every replica, renderPasses() renders every replica's share of the
// frame (64x64 tiles round-robin) or of the samples (sampleCount = p * N + rank), readAverage() combines the accumulators on
// rank 0 through the C ABI's vt_group (NCCL over NVLink, or the library's peer-memory kernel). Single host thread, like the
// reference's GL thread: every call only enqueues work on the replicas' streams.
class RendererGroup {
public:
    enum Mode { MODE_TILES = VT_PART_TILES, MODE_SAMPLES = VT_PART_SAMPLES };
    RendererGroup();
    ~RendererGroup();
    bool initialize(const std::vector<int>& cudaDevices, Mode mode);
    size_t size() const { return m_renderers.size(); }
    Renderer& renderer(size_t rank) { return *m_renderers[rank]; }
    template <class F> void forEach(F f) { for (size_t i = 0; i < m_renderers.size(); ++i) f(*m_renderers[i]); }   // any Renderer call, on every replica
    void resizeFrame(int width, int height);
    void loadVoxFile(const std::string& file);
    void loadMeshAtResolution(const std::string& file, int resolution);
    void setVoxelData(const vtm::V3i& resolution, const std::vector<int32_t>& voxelMaterials, const std::vector<float>& materialData,
                      const std::vector<int32_t>& emissiveVoxelIndices);
    void updateRenderSettings(const RenderSettings& settings);
    void requestAction(float x, float y, float dx, float dy, Action::PICKING_ACTION action, bool restartAccumulation);   // actions.cpp:5-18, on every replica
    void resetRender();
    void renderPasses(int nPasses);
    bool beginCombine();                       // asynchronous: rendering may continue while the exchange is in flight
    bool endCombine(float* rgbaOut);           // W*H RGBA float32, GL orientation
    bool readAverage(float* rgbaOut) { return beginCombine() && endCombine(rgbaOut); }
    vt_group* group() const { return m_group; }
    const std::string& getStatus() const { return m_status; }
private:
    std::vector<Renderer*> m_renderers;
    vt_group* m_group;
    std::string m_status;
};

// ---- tools/tool.h:5-13, toolAddRemoveVoxel, toolFocalDistance -------------------------------------------------------------
struct MouseEvent { int x, y; int buttons; int modifiers; };      // stands in for QMouseEvent
struct KeyEvent { int key; };
struct WidgetSize { int width, height; };
class Tool {
public:
    virtual ~Tool() {}
    virtual bool mousePressEvent(const MouseEvent*, WidgetSize) { return false; }
    virtual bool mouseReleaseEvent(const MouseEvent*, WidgetSize) { return false; }
    virtual bool mouseMoveEvent(const MouseEvent*, WidgetSize) { return false; }
    virtual bool keyPressEvent(const KeyEvent*) { return false; }
};
class ToolAddRemoveVoxel : public Tool {
public:
    explicit ToolAddRemoveVoxel(Renderer& r) : m_renderer(r) { m_lastPos[0] = m_lastPos[1] = 0; }
    bool mousePressEvent(const MouseEvent* event, WidgetSize widgetDimensions) override;
    bool mouseMoveEvent(const MouseEvent* event, WidgetSize widgetDimensions) override;
private:
    Renderer& m_renderer;
    int m_lastPos[2];
};
class ToolFocalDistance : public Tool {
public:
    explicit ToolFocalDistance(Renderer& r) : m_renderer(r) {}
    bool mousePressEvent(const MouseEvent* event, WidgetSize widgetDimensions) override;
private:
    Renderer& m_renderer;
};

// ---- ui/glwidget.h: the widget that drives the renderer, without Qt ----------------------------------
// Same members, slots and event routing as GLWidget (ui/glwidget.cpp:22-410); Qt's update() -> paintGL() round trip is an
// explicit flag that pump() drains, so a script (runScript) or a server loop can play real event traffic against Renderer.
class HeadlessWidget {
public:
    enum ResolutionMode { RM_FIXED, RM_LONGEST_AXIS, RM_MATCH_WINDOW };          // ui/renderpropertiesui.h
    enum ToolAction { ACTION_SELECT_FOCAL_POINT, ACTION_EDIT_VOXELS };            // ui/mainwindow.h
    explicit HeadlessWidget(Renderer& renderer);
    ~HeadlessWidget();
    // QGLWidget side
    void resizeGL(int width, int height);
    bool paintGL();                                   // true when another repaint was scheduled (samples pending)
    int pump(int maxPaints);                          // runs paintGL while an update is pending; returns the number of paints
    void mousePressEvent(const MouseEvent& e);
    void mouseMoveEvent(const MouseEvent& e);
    void mouseReleaseEvent(const MouseEvent& e);
    void keyPressEvent(const KeyEvent& e);
    // slots (glwidget.cpp:225-410)
    void cameraFStopChanged(float fstop);
    void cameraFocalLengthChanged(float length);
    void cameraLensModelChanged(int model);
    void cameraControllerChanged(const std::string& mode);
    void onPathtracerMaxSamplesChanged(int value);
    void onPathtracerMaxPathBouncesChanged(int value);
    void onWireframeOpacityChanged(int value);
    void onWireframeThicknessChanged(int value);
    void loadMesh(const std::string& file);
    size_t loadVoxFile(const std::string& file);      // returns the number of materials announced (materialCreated signals)
    void saveImage(const std::string& file);
    void onResolutionSettingsChanged(ResolutionMode mode, int axis1, int axis2);
    void onActionTriggered(int action, bool triggered);
    void onBackgroundColorChangedConstant(const vtm::V3f& color);
    void onBackgroundColorChangedGradientFrom(const vtm::V3f& color);
    void onBackgroundColorChangedGradientTo(const vtm::V3f& color);
    void onBackgroundColorChangedImage(const std::string& path);
    void onBackgroundImageRotationChanged(int rotation);
    void onBeginUserInteraction();
    void onEndUserInteraction();
    void onMaterialColorChanged(unsigned int dataOffset, const float rgb[3]);
    void onMaterialValueChanged(unsigned int dataOffset, float value);
    // script player: one command per line (see headless.cpp); returns false and fills `error` at the first bad line
    bool runScript(const std::string& script, std::string& error);
    bool updatePending() const { return m_updatePending; }
    int width() const { return m_width; }
    int height() const { return m_height; }
    unsigned long paints() const { return m_paints; }
private:
    void update() { m_updatePending = true; }
    void resizeRender(int renderW, int renderH, int windowW, int windowH);
    Renderer& m_renderer;
    ResolutionMode m_resolutionMode;
    int m_resolutionLongestAxis;
    Tool* m_activeTool;
    unsigned int m_activeUserDialogs;
    int m_lastPos[2];
    int m_lastMouseButtons;
    int m_width, m_height;
    bool m_updatePending;
    unsigned long m_paints;
};

