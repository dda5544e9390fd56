// headless.cpp -- HeadlessWidget: the reference's GLWidget (ui/glwidget.cpp) without Qt, and a script player over it.
//
// GLWidget is the only caller of the Renderer class API: it routes mouse / key events to the active tool
// first and to the renderer second, turns UI slots into RenderSettings edits, and repaints while samples are pending.
// This file keeps that routing verbatim in behaviour (which handler sees an event, what resets the accumulation, what
// schedules a repaint) so that event traffic recorded against the reference UI can be replayed against this build.
// Qt's `update()` -> event loop -> `paintGL()` becomes a flag drained by pump().
#include "vt_host.h"

#include <algorithm>
#include <cstdlib>
#include <sstream>

using vtm::V3f;

HeadlessWidget::HeadlessWidget(Renderer& renderer)                               // glwidget.cpp:22-31
    : m_renderer(renderer), m_resolutionMode(RM_MATCH_WINDOW), m_resolutionLongestAxis(1024), m_activeTool(NULL),
      m_activeUserDialogs(0U), m_lastMouseButtons(0), m_width(400), m_height(400), m_updatePending(true), m_paints(0)
{
    m_lastPos[0] = m_lastPos[1] = 0;                                             // sizeHint 400x400 (:42-45)
}

HeadlessWidget::~HeadlessWidget() { delete m_activeTool; }

void HeadlessWidget::resizeRender(int renderW, int renderH, int windowW, int windowH)      // glwidget.cpp:62-141
{
    int resW = renderW, resH = renderH;
    const float windowAR = (float)windowW / windowH;
    int viewportX = 0, viewportY = 0, viewportW = windowW, viewportH = windowH;
    const float renderAR = (float)renderW / renderH;
    switch (m_resolutionMode) {
    case RM_FIXED:
        if (windowAR >= renderAR) {                                              // pillar box
            viewportY = 0; viewportH = windowH;
            viewportW = (int)(windowH * renderAR);
            viewportX = (windowW - viewportW) / 2;
        } else {                                                                 // letter box
            viewportX = 0; viewportW = windowW;
            viewportH = (int)(windowW / renderAR);
            viewportY = (windowH - viewportH) / 2;
        }
        break;
    case RM_LONGEST_AXIS:
        if (windowW >= windowH) { resW = m_resolutionLongestAxis; resH = (int)(resW / windowAR); }
        else { resH = m_resolutionLongestAxis; resW = (int)(resH * windowAR); }
        break;
    case RM_MATCH_WINDOW:
        resW = windowW; resH = windowH;
        break;
    }
    m_renderer.resizeFrame(resW, resH, viewportX, viewportY, viewportW, viewportH);
    update();
}

void HeadlessWidget::resizeGL(int width, int height)                             // glwidget.cpp:142-148
{
    m_width = width; m_height = height;
    resizeRender(m_renderer.renderSettings().m_imageResolution.x, m_renderer.renderSettings().m_imageResolution.y, width, height);
}

bool HeadlessWidget::paintGL()                                                   // glwidget.cpp:150-159
{
    m_updatePending = false;
    ++m_paints;
    const bool pauseRendering = m_activeUserDialogs > 0;
    if (!pauseRendering && m_renderer.render() == Renderer::RR_SAMPLES_PENDING) update();
    return m_updatePending;
}

int HeadlessWidget::pump(int maxPaints)
{
    int n = 0;
    while (m_updatePending && n < maxPaints) { paintGL(); ++n; }
    return n;
}

void HeadlessWidget::mousePressEvent(const MouseEvent& e)                        // glwidget.cpp:161-169
{
    const WidgetSize size = { m_width, m_height };
    if (m_activeTool != NULL && m_activeTool->mousePressEvent(&e, size)) update();
}

void HeadlessWidget::mouseMoveEvent(const MouseEvent& e)                         // glwidget.cpp:171-192
{
    const WidgetSize size = { m_width, m_height };
    if (m_activeTool != NULL && m_activeTool->mouseMoveEvent(&e, size)) { update(); return; }   // note: m_lastPos is NOT advanced
    const int dx = e.x - m_lastPos[0], dy = e.y - m_lastPos[1];
    if (m_renderer.onMouseMove(dx, dy, e.buttons)) update();
    m_lastPos[0] = e.x; m_lastPos[1] = e.y;
    m_lastMouseButtons = e.buttons;
}

void HeadlessWidget::mouseReleaseEvent(const MouseEvent& e)                      // glwidget.cpp:194-206
{
    const WidgetSize size = { m_width, m_height };
    if (m_activeTool != NULL && m_activeTool->mouseReleaseEvent(&e, size)) { update(); return; }
    m_lastPos[0] = e.x; m_lastPos[1] = e.y;
    m_lastMouseButtons = e.buttons;
}

void HeadlessWidget::keyPressEvent(const KeyEvent& e)                            // glwidget.cpp:208-223
{
    if (m_activeTool != NULL && m_activeTool->keyPressEvent(&e)) { update(); return; }
    if (m_renderer.onKeyPress(e.key)) update();
}

// ---- slots ---------------------------------------------------------------------------------------------------------------
void HeadlessWidget::cameraFStopChanged(float fstop) { m_renderer.camera().setFStop(fstop); m_renderer.resetRender(); update(); }              // :225-230
void HeadlessWidget::cameraFocalLengthChanged(float length) { m_renderer.camera().setFocalLength(length); m_renderer.resetRender(); update(); } // :232-237
void HeadlessWidget::cameraLensModelChanged(int model)                                                                                         // :239-244
{
    m_renderer.camera().setLensModel((CameraParameters::CameraLensModel)model); m_renderer.resetRender(); update();
}
void HeadlessWidget::cameraControllerChanged(const std::string& mode)                                                                          // :245-255
{
    m_renderer.camera().setCameraController(mode == "orbit" ? Camera::CCM_ORBIT : Camera::CCM_FLY);
}
void HeadlessWidget::onPathtracerMaxSamplesChanged(int value) { m_renderer.renderSettings().m_pathtracerMaxSamples = value; m_renderer.updateRenderSettings(); update(); }
void HeadlessWidget::onPathtracerMaxPathBouncesChanged(int value) { m_renderer.renderSettings().m_pathtracerMaxNumBounces = value; m_renderer.updateRenderSettings(); update(); }
void HeadlessWidget::onWireframeOpacityChanged(int value) { m_renderer.renderSettings().m_wireframeOpacity = (float)value / 100; m_renderer.updateRenderSettings(); update(); }
void HeadlessWidget::onWireframeThicknessChanged(int value) { m_renderer.renderSettings().m_wireframeThickness = (float)value / 1000; m_renderer.updateRenderSettings(); update(); }
void HeadlessWidget::loadMesh(const std::string& file) { m_renderer.loadMesh(file); }                                                          // :285-288
size_t HeadlessWidget::loadVoxFile(const std::string& file)                                                                                    // :290-299
{
    m_renderer.loadVoxFile(file);
    return m_renderer.getMaterials().size();          // one materialCreated signal per entry in the reference
}
void HeadlessWidget::saveImage(const std::string& file) { m_renderer.saveImage(file); }                                                        // :301-304
void HeadlessWidget::onResolutionSettingsChanged(ResolutionMode mode, int axis1, int axis2)                                                    // :314-321
{
    m_resolutionMode = mode; m_resolutionLongestAxis = axis1;
    resizeRender(axis1, axis2, m_width, m_height);
}
void HeadlessWidget::onActionTriggered(int action, bool triggered)                                                                             // :323-348
{
    if (!triggered) { delete m_activeTool; m_activeTool = NULL; return; }
    switch (action) {
    case ACTION_SELECT_FOCAL_POINT: delete m_activeTool; m_activeTool = new ToolFocalDistance(m_renderer); break;   // (the reference leaks the old tool)
    case ACTION_EDIT_VOXELS: delete m_activeTool; m_activeTool = new ToolAddRemoveVoxel(m_renderer); break;
    default: return;
    }
}
void HeadlessWidget::onBackgroundColorChangedConstant(const V3f& c)                                                                            // :350-357
{
    m_renderer.renderSettings().m_backgroundImage = "";
    m_renderer.renderSettings().m_backgroundColor[0] = c; m_renderer.renderSettings().m_backgroundColor[1] = c;
    m_renderer.updateRenderSettings(); update();
}
void HeadlessWidget::onBackgroundColorChangedGradientFrom(const V3f& c)                                                                        // :358-364
{
    m_renderer.renderSettings().m_backgroundImage = ""; m_renderer.renderSettings().m_backgroundColor[0] = c;
    m_renderer.updateRenderSettings(); update();
}
void HeadlessWidget::onBackgroundColorChangedGradientTo(const V3f& c)                                                                          // :365-371
{
    m_renderer.renderSettings().m_backgroundImage = ""; m_renderer.renderSettings().m_backgroundColor[1] = c;
    m_renderer.updateRenderSettings(); update();
}
void HeadlessWidget::onBackgroundColorChangedImage(const std::string& path)                                                                    // :372-377
{
    m_renderer.renderSettings().m_backgroundImage = path; m_renderer.updateRenderSettings(); update();
}
void HeadlessWidget::onBackgroundImageRotationChanged(int rotation)                                                                            // :378-383
{
    m_renderer.renderSettings().m_backgroundRotationDegrees = rotation; m_renderer.updateRenderSettings(); update();
}
void HeadlessWidget::onBeginUserInteraction() { m_activeUserDialogs++; }                                                                       // :389-392
void HeadlessWidget::onEndUserInteraction() { m_activeUserDialogs = std::max(1U, m_activeUserDialogs) - 1; }                                   // :394-397
void HeadlessWidget::onMaterialColorChanged(unsigned int dataOffset, const float rgb[3])                                                       // :400-406
{
    m_renderer.updateMaterialColor(dataOffset, rgb); m_renderer.resetRender(); update();
}
void HeadlessWidget::onMaterialValueChanged(unsigned int dataOffset, float value)                                                              // :407-412
{
    m_renderer.updateMaterialValue(dataOffset, value); m_renderer.resetRender(); update();
}

// ---- script player -------------------------------------------------------------------------------------------------------
// One command per line, '#' starts a comment. Coordinates are widget pixels (origin top-left), buttons / modifiers / keys are
// the Qt values (vtinput) or the names below.
//   window W H                       resizeGL
//   resolution fixed|longest|window A1 [A2]
//   vox FILE | mesh FILE             loadVoxFile / loadMesh
//   tool none|focal|edit             onActionTriggered
//   press|move|release X Y [BUTTONS [MODIFIERS]]     BUTTONS: left,right,middle,none or a number; MODIFIERS: ctrl,none or a number
//   key space|f|w|a|s|d|CODE
//   fstop V | focal_length V | lens N | controller orbit|fly
//   max_samples N | max_bounces N | wireframe_opacity N | wireframe_thickness N
//   background constant|from|to R G B | background image FILE | background rotation DEG
//   material_color OFFSET R G B | material_value OFFSET V
//   dialog begin|end
//   paint [N]                        at most N repaints while an update is pending (default: until rendering finishes, cap 1e6)
//   save FILE
static int parseButtons(const std::string& s)
{
    if (s == "left") return vtinput::LeftButton; if (s == "right") return vtinput::RightButton; if (s == "middle") return vtinput::MiddleButton;
    if (s == "none" || s.empty()) return 0;
    return atoi(s.c_str());
}
static int parseModifiers(const std::string& s) { if (s == "ctrl") return vtinput::ControlModifier; if (s == "none" || s.empty()) return 0; return (int)strtol(s.c_str(), NULL, 0); }
static int parseKey(const std::string& s)
{
    if (s == "space") return vtinput::Key_Space;
    if (s.size() == 1 && isalpha((unsigned char)s[0])) return toupper((unsigned char)s[0]);      // Qt::Key_A..Z are the upper-case ASCII codes
    return (int)strtol(s.c_str(), NULL, 0);
}

bool HeadlessWidget::runScript(const std::string& script, std::string& error)
{
    std::istringstream in(script);
    std::string line;
    int lineNo = 0;
    while (std::getline(in, line)) {
        ++lineNo;
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        std::istringstream ls(line);
        std::string cmd;
        if (!(ls >> cmd)) continue;
        bool ok = true;
        if (cmd == "window") { int w = 0, h = 0; ok = bool(ls >> w >> h) && w > 0 && h > 0; if (ok) resizeGL(w, h); }
        else if (cmd == "resolution") {
            std::string m; int a1 = 0, a2 = 0; ok = bool(ls >> m); ls >> a1 >> a2;
            if (ok && m == "fixed") { ok = a1 > 0 && a2 > 0; if (ok) onResolutionSettingsChanged(RM_FIXED, a1, a2); }
            else if (ok && m == "longest") { ok = a1 > 0; if (ok) onResolutionSettingsChanged(RM_LONGEST_AXIS, a1, a2 > 0 ? a2 : a1); }
            else if (ok && m == "window") onResolutionSettingsChanged(RM_MATCH_WINDOW, m_width, m_height);
            else ok = false;
        }
        else if (cmd == "vox" || cmd == "mesh") { std::string f; ok = bool(ls >> f); if (ok) { if (cmd == "vox") loadVoxFile(f); else loadMesh(f); update(); } }
        else if (cmd == "tool") {
            std::string t; ok = bool(ls >> t);
            if (ok && t == "none") onActionTriggered(0, false);
            else if (ok && t == "focal") onActionTriggered(ACTION_SELECT_FOCAL_POINT, true);
            else if (ok && t == "edit") onActionTriggered(ACTION_EDIT_VOXELS, true);
            else ok = false;
        }
        else if (cmd == "press" || cmd == "move" || cmd == "release") {
            MouseEvent e = { 0, 0, 0, 0 }; std::string b, m; ok = bool(ls >> e.x >> e.y); ls >> b >> m;
            e.buttons = parseButtons(b); e.modifiers = parseModifiers(m);
            if (ok) { if (cmd == "press") mousePressEvent(e); else if (cmd == "move") mouseMoveEvent(e); else mouseReleaseEvent(e); }
        }
        else if (cmd == "key") { std::string k; ok = bool(ls >> k); if (ok) { KeyEvent e = { parseKey(k) }; keyPressEvent(e); } }
        else if (cmd == "fstop") { float v; ok = bool(ls >> v); if (ok) cameraFStopChanged(v); }
        else if (cmd == "focal_length") { float v; ok = bool(ls >> v); if (ok) cameraFocalLengthChanged(v); }
        else if (cmd == "lens") { int v; ok = bool(ls >> v) && v >= 0 && v <= 2; if (ok) cameraLensModelChanged(v); }
        else if (cmd == "controller") { std::string m; ok = bool(ls >> m) && (m == "orbit" || m == "fly"); if (ok) cameraControllerChanged(m); }
        else if (cmd == "max_samples") { int v; ok = bool(ls >> v); if (ok) onPathtracerMaxSamplesChanged(v); }
        else if (cmd == "max_bounces") { int v; ok = bool(ls >> v); if (ok) onPathtracerMaxPathBouncesChanged(v); }
        else if (cmd == "wireframe_opacity") { int v; ok = bool(ls >> v); if (ok) onWireframeOpacityChanged(v); }
        else if (cmd == "wireframe_thickness") { int v; ok = bool(ls >> v); if (ok) onWireframeThicknessChanged(v); }
        else if (cmd == "background") {
            std::string m; ok = bool(ls >> m);
            if (ok && (m == "constant" || m == "from" || m == "to")) {
                V3f c; ok = bool(ls >> c.x >> c.y >> c.z);
                if (ok) { if (m == "constant") onBackgroundColorChangedConstant(c); else if (m == "from") onBackgroundColorChangedGradientFrom(c); else onBackgroundColorChangedGradientTo(c); }
            }
            else if (ok && m == "image") { std::string f; ok = bool(ls >> f); if (ok) onBackgroundColorChangedImage(f); }
            else if (ok && m == "rotation") { int d; ok = bool(ls >> d); if (ok) onBackgroundImageRotationChanged(d); }
            else ok = false;
        }
        else if (cmd == "material_color") { unsigned int off; float c[3]; ok = bool(ls >> off >> c[0] >> c[1] >> c[2]); if (ok) onMaterialColorChanged(off, c); }
        else if (cmd == "material_value") { unsigned int off; float v; ok = bool(ls >> off >> v); if (ok) onMaterialValueChanged(off, v); }
        else if (cmd == "dialog") { std::string m; ok = bool(ls >> m) && (m == "begin" || m == "end"); if (ok) { if (m == "begin") onBeginUserInteraction(); else onEndUserInteraction(); } }
        else if (cmd == "paint") { int n = 1000000; ls >> n; pump(n); }
        else if (cmd == "save") { std::string f; ok = bool(ls >> f); if (ok) saveImage(f); }
        else ok = false;
        if (!ok) {
            std::ostringstream msg; msg << "line " << lineNo << ": cannot run '" << line << "'";
            error = msg.str();
            return false;
        }
    }
    return true;
}
