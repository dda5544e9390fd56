// camera.cpp -- vector helpers + CameraParameters / Camera / controllers.
// Behaviour follows camera/cameraParameters.cpp, camera/camera.cpp, camera/cameraController.cpp,
// camera/orbitCameraController.cpp and camera/flyCameraController.cpp of the reference; the Qt
// button / key enums are replaced by vtinput::*.
#include "vt_host.h"

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>

namespace vtm {

float V3f::length() const { return std::sqrt(x * x + y * y + z * z); }
V3f V3f::normalized() const
{
    const float l = length();
    if (l == 0.0f) return V3f(0.0f);
    return V3f(x / l, y / l, z / l);
}
void Box3f::makeEmpty() { min = V3f(FLT_MAX); max = V3f(-FLT_MAX); }
void Box3f::extendBy(const V3f& p)
{
    for (int i = 0; i < 3; ++i) { if (p[i] < min[i]) min[i] = p[i]; if (p[i] > max[i]) max[i] = p[i]; }
}
int Box3f::majorAxis() const
{
    const V3f s = size();
    int major = 0;
    if (s[1] > s[major]) major = 1;
    if (s[2] > s[major]) major = 2;
    return major;
}
void M44f::makeIdentity()
{
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) x[r][c] = (r == c) ? 1.0f : 0.0f;
}
M44f M44f::operator*(const M44f& o) const
{
    M44f out;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += (double)x[r][k] * o.x[k][c];
        out.x[r][c] = (float)s;
    }
    return out;
}
M44f M44f::inverse() const
{
    // Gauss-Jordan with partial pivoting in double; a singular matrix returns identity (Imath's behaviour)
    double a[4][8];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = x[r][c]; a[r][c + 4] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r) if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) return M44f();
        if (piv != col) for (int c = 0; c < 8; ++c) std::swap(a[piv][c], a[col][c]);
        const double d = a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] /= d;
        for (int r = 0; r < 4; ++r) if (r != col) {
            const double f = a[r][col];
            if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    M44f out;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out.x[r][c] = (float)a[r][c + 4];
    return out;
}
} // namespace vtm

using namespace vtm;

// ---- CameraParameters (cameraParameters.cpp:4-184) -------------------------------------------------------
float CameraParameters::FILM_SIZE_35MM = 36.0f;

CameraParameters::CameraParameters()
    : m_target(0, 0, 0), m_eye(0, 0, -1), m_fovY(0), m_near(0.1f), m_far(10000), m_focalDistance(100),
      m_lensRadius(0), m_filmSize(36.0f),
      m_lensModel(CLM_PINHOLE)      // never initialised in the reference; pinned to pinhole here
{
    setFocalLength(50);
}
void CameraParameters::lookAt(const V3f& target) { m_target = target; }
float CameraParameters::distanceToTarget() const { return (m_target - m_eye).length(); }
void CameraParameters::setDistanceFromTarget(float distance) { m_eye = m_target - forwardUnitVector() * distance; }
void CameraParameters::getBasis(V3f& fwd, V3f& right, V3f& up) const
{
    fwd = (m_target - m_eye).normalized();
    right = V3f(0, 1, 0).cross(fwd).normalized();
    up = fwd.cross(right);
}
V3f CameraParameters::forwardUnitVector() const { V3f f, r, u; getBasis(f, r, u); return f; }
V3f CameraParameters::rightUnitVector() const { V3f f, r, u; getBasis(f, r, u); return r; }
V3f CameraParameters::upUnitVector() const { V3f f, r, u; getBasis(f, r, u); return u; }
float CameraParameters::rotationTheta() const { return std::acos(forwardUnitVector().y); }
float CameraParameters::rotationPhi() const { const V3f f = forwardUnitVector(); return std::atan2(f.z, f.x); }

static V3f sphericalDirection(float theta, float phi)
{
    const float st = std::sin(theta);
    return V3f(st * std::cos(phi), std::cos(theta), st * std::sin(phi));
}
void CameraParameters::orbitAroundTarget(float theta, float phi)
{
    const float r = distanceToTarget();
    m_eye = m_target - r * sphericalDirection(theta, phi);
}
void CameraParameters::orbitAroundEye(float theta, float phi)
{
    const float r = distanceToTarget();
    m_target = m_eye + r * sphericalDirection(theta, phi);
}
void CameraParameters::setEyeTarget(const V3f& eye, const V3f& target) { m_eye = eye; m_target = target; }
void CameraParameters::setFovY(float fov) { m_fovY = fov; }
void CameraParameters::setNearDistance(float d) { m_near = std::max(0.0f, d); }
void CameraParameters::setFarDistance(float d) { m_far = std::max(m_near, d); }
float CameraParameters::focalLength() const { return m_filmSize.y / (2.0f * std::tan(0.5f * m_fovY)); }
void CameraParameters::setFocalLength(float focalLength) { m_fovY = std::atan2(m_filmSize.y * 0.5f, focalLength) * 2.0f; }
void CameraParameters::setFocalDistance(float distance) { m_focalDistance = std::max(0.0f, distance); }
void CameraParameters::setFilmSize(float filmW, float filmH)
{
    m_filmSize = V2f(std::max(0.0f, filmW), std::max(0.0f, filmH));
    m_fovY = std::atan2(m_filmSize.y * 0.5f, focalLength()) * 2.0f;    // keeps the vertical field of view
}
void CameraParameters::setLensRadius(float radius) { m_lensRadius = std::max(0.0f, radius); }
void CameraParameters::setFStop(float fstop) { m_lensRadius = (focalLength() / std::max(1e-4f, fstop)) * 0.5f; }
void CameraParameters::setLensModel(CameraLensModel model) { m_lensModel = model; }

// ---- controllers -----------------------------------------------------------------------------------------
void CameraController::lookAt(const V3f& target) { if (m_parameters) m_parameters->lookAt(target); }
void CameraController::setDistanceFromTarget(float d) { if (m_parameters) m_parameters->setDistanceFromTarget(d); }
void CameraController::orbitAroundTarget(float theta, float phi) { if (m_parameters) m_parameters->orbitAroundTarget(theta, phi); }
void CameraController::focusOnBounds(const Box3f& bounds)          // cameraController.cpp:16-23
{
    const V3f fwd = m_parameters->forwardUnitVector();
    const V3f center = bounds.center();
    const float distance = bounds.size().length() / (2.0f * std::tan(m_parameters->fovY() / 2));
    m_parameters->setEyeTarget(center - fwd * distance, center);
}

bool OrbitCameraController::onMouseMove(float dx, float dy, int buttons)   // orbitCameraController.cpp:15-41
{
    if (!m_parameters) return false;
    if (buttons & vtinput::RightButton) {
        // float * double(M_PI) evaluated in double, as the reference's expression is
        const float theta = (float)((double)dy * M_PI * 1.0 + (double)m_parameters->rotationTheta());
        const float phi = (float)((double)(-dx * 2.0f) * M_PI * 1.0 + (double)m_parameters->rotationPhi());
        m_parameters->orbitAroundTarget(theta, phi);
        return true;
    }
    if (buttons & vtinput::MiddleButton) {
        const float speed = 1.05f;
        const float d = m_parameters->distanceToTarget();
        m_parameters->setDistanceFromTarget(dy > 0 ? d * speed : d / speed);
        return true;
    }
    return false;
}
bool OrbitCameraController::onKeyPress(int) { return false; }

bool FlyCameraController::onMouseMove(float dx, float dy, int buttons)     // flyCameraController.cpp:15-40
{
    if (!m_parameters) return false;
    if (buttons & vtinput::RightButton) {
        const float theta = (float)((double)dy * M_PI * 1.0 + (double)m_parameters->rotationTheta());
        const float phi = (float)((double)(-dx * 2.0f) * M_PI * 1.0 + (double)m_parameters->rotationPhi());
        m_parameters->orbitAroundEye(theta, phi);
        return true;
    }
    if (buttons & vtinput::MiddleButton) {
        const V3f up = m_parameters->upUnitVector() * (dy * 100.0f);
        m_parameters->setEyeTarget(m_parameters->eye() + up, m_parameters->target() + up);
        return true;
    }
    return false;
}
bool FlyCameraController::onKeyPress(int key)                                // flyCameraController.cpp:42-73
{
    const float speed = 100;
    V3f step;
    switch (key) {
    case vtinput::Key_W: step = m_parameters->forwardUnitVector() * speed; break;
    case vtinput::Key_S: step = m_parameters->forwardUnitVector() * -speed; break;
    case vtinput::Key_D: step = m_parameters->rightUnitVector() * speed; break;
    case vtinput::Key_A: step = m_parameters->rightUnitVector() * -speed; break;
    default: return false;
    }
    m_parameters->setEyeTarget(m_parameters->eye() + step, m_parameters->target() + step);
    return true;
}

// ---- Camera (camera.cpp:6-38) ---------------------------------------------------------------------------------
Camera::Camera() : m_controller(new OrbitCameraController(&m_parameters)), m_controllerMode(CCM_ORBIT) {}
Camera::~Camera() { delete m_controller; }
void Camera::setLensModel(CameraParameters::CameraLensModel m) { m_parameters.setLensModel(m); }
void Camera::setFocalLength(float v) { m_parameters.setFocalLength(v); }
void Camera::setFocalDistance(float v) { m_parameters.setFocalDistance(v); }
void Camera::setFilmSize(float w, float h) { m_parameters.setFilmSize(w, h); }
void Camera::setLensRadius(float v) { m_parameters.setLensRadius(v); }
void Camera::setFStop(float v) { m_parameters.setFStop(v); }
void Camera::setCameraController(CameraControllerMode mode)
{
    if (m_controllerMode == mode) return;
    m_controllerMode = mode;
    delete m_controller;                       // the reference leaks the previous controller here
    if (mode == CCM_ORBIT) m_controller = new OrbitCameraController(&m_parameters);
    else m_controller = new FlyCameraController(&m_parameters);
}
