// tetsim_napi.cc -- N-API shim: exposes include/tetsim_b200.h to Node one call per entry point.
// SOURCE ONLY in this repository: the build image has no Node, no node_api.h and no JS engine, so
// this file is not compiled or tested here.  Build where Node is available:
//   g++ -std=c++17 -shared -fPIC -I$(node -p "require('node:process').execPath + '/../../include/node'") \
//       -I../../include tetsim_napi.cc -L.. -ltetsim_b200 -Wl,-rpath,'$ORIGIN/..' -o tetsim_napi.node
// Typed arrays are passed as-is (Float32Array / Int32Array backing stores); numbers are doubles.
#include <node_api.h>

#include <cstring>
#include <string>

#include "tetsim_b200.h"

#define NAPI_OK(call)                                                                  \
    do {                                                                               \
        if ((call) != napi_ok) { napi_throw_error(env, nullptr, "N-API failure: " #call); return nullptr; } \
    } while (0)

static napi_value throw_tetsim(napi_env env, int rc) {
    // the reference's "error string or null" convention (MultiTargetGPUComputationRenderer.js:178-190)
    // surfaces as a thrown Error carrying the library's message
    std::string msg = std::string("tetsim error ") + std::to_string(rc) + ": " + tetsim_last_error();
    napi_throw_error(env, nullptr, msg.c_str());
    return nullptr;
}

static bool get_f64(napi_env env, napi_value obj, const char *key, double *out) {
    napi_value v;
    bool has = false;
    if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return false;
    return napi_get_named_property(env, obj, key, &v) == napi_ok && napi_get_value_double(env, v, out) == napi_ok;
}

// physicsParams object (src/main.js:22-36) -> TetSimParams; missing keys keep the defaults
static void read_params(napi_env env, napi_value obj, TetSimParams *p) {
    tetsim_default_params(p);
    napi_valuetype t;
    if (napi_typeof(env, obj, &t) != napi_ok || t != napi_object) return;
    get_f64(env, obj, "gravity", &p->gravity);
    get_f64(env, obj, "friction", &p->friction);
    get_f64(env, obj, "density", &p->density);
    get_f64(env, obj, "devCompliance", &p->devCompliance);
    get_f64(env, obj, "volCompliance", &p->volCompliance);
    napi_value wb;
    bool has = false;
    if (napi_has_named_property(env, obj, "worldBounds", &has) == napi_ok && has &&
        napi_get_named_property(env, obj, "worldBounds", &wb) == napi_ok)
        for (uint32_t i = 0; i < 6; i++) {
            napi_value e;
            if (napi_get_element(env, wb, i, &e) == napi_ok) napi_get_value_double(env, e, &p->worldBounds[i]);
        }
}

template <class T>
static T *typed(napi_env env, napi_value v, size_t *len) {
    napi_typedarray_type ty;
    void *data = nullptr;
    napi_value ab;
    size_t off;
    if (napi_get_typedarray_info(env, v, &ty, len, &data, &ab, &off) != napi_ok) return nullptr;
    return static_cast<T *>(data);
}

static tetsim_t *unwrap(napi_env env, napi_value v) {
    void *p = nullptr;
    napi_get_value_external(env, v, &p);
    return static_cast<tetsim_t *>(p);
}

// create(vertices: Float32Array, tetIds: Int32Array, physicsParams, options{solver,arithmetic,iters,...}) -> handle
static napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    size_t nv = 0, nt = 0;
    float *verts = typed<float>(env, argv[0], &nv);
    int32_t *ids = typed<int32_t>(env, argv[1], &nt);
    TetSimParams prm;
    read_params(env, argv[2], &prm);
    TetSimOptions opt;
    tetsim_default_options(&opt);
    double d;
    if (argc > 3) {
        if (get_f64(env, argv[3], "solver", &d)) opt.solver = (int32_t)d;
        if (get_f64(env, argv[3], "arithmetic", &d)) opt.arithmetic = (int32_t)d;
        if (get_f64(env, argv[3], "iters", &d)) opt.iters = (int32_t)d;
        if (get_f64(env, argv[3], "deterministic", &d)) opt.deterministic = (int32_t)d;
        if (get_f64(env, argv[3], "referenceTableBug", &d)) opt.referenceTableBug = (int32_t)d;
        if (get_f64(env, argv[3], "clusterSize", &d)) opt.clusterSize = (int32_t)d;
        if (get_f64(env, argv[3], "device", &d)) opt.device = (int32_t)d;
    }
    tetsim_t *h = nullptr;
    int rc = tetsim_create(verts, (int32_t)(nv / 3), ids, (int32_t)(nt / 4), &prm, &opt, &h);
    if (rc != TETSIM_OK) return throw_tetsim(env, rc);
    napi_value ext;
    NAPI_OK(napi_create_external(env, h, [](napi_env, void *data, void *) { tetsim_destroy(static_cast<tetsim_t *>(data)); },
                                 nullptr, &ext));
    return ext;
}

// simulate(handle, dt, physicsParams)          == softBody.simulate(dt, physicsParams)
static napi_value Simulate(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double dt = 0;
    napi_get_value_double(env, argv[1], &dt);
    TetSimParams prm;
    read_params(env, argv[2], &prm);
    int rc = tetsim_simulate(unwrap(env, argv[0]), dt, &prm);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// step(handle, frameDt, numSubsteps, physicsParams)   == the loop at src/main.js:79-84
static napi_value Step(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double frameDt = 0, n = 1;
    napi_get_value_double(env, argv[1], &frameDt);
    napi_get_value_double(env, argv[2], &n);
    TetSimParams prm;
    read_params(env, argv[3], &prm);
    int rc = tetsim_step(unwrap(env, argv[0]), frameDt, (int32_t)n, &prm);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// readPositions(handle, out: Float32Array) / readVelocities / readPrevPositions
template <int (*FN)(tetsim_t *, float *)>
static napi_value Read3(napi_env env, napi_callback_info info) {
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    size_t n = 0;
    float *out = typed<float>(env, argv[1], &n);
    int rc = FN(unwrap(env, argv[0]), out);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// startGrab(handle, x, y, z) -> grabId ; moveGrabbed(handle, x, y, z) ; endGrab(handle)
static napi_value StartGrab(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double p[3];
    for (int i = 0; i < 3; i++) napi_get_value_double(env, argv[1 + i], &p[i]);
    int32_t id = -1;
    int rc = tetsim_start_grab(unwrap(env, argv[0]), p, &id);
    if (rc != TETSIM_OK) return throw_tetsim(env, rc);
    napi_value out;
    NAPI_OK(napi_create_int32(env, id, &out));
    return out;
}
static napi_value MoveGrabbed(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double p[3];
    for (int i = 0; i < 3; i++) napi_get_value_double(env, argv[1 + i], &p[i]);
    int rc = tetsim_move_grabbed(unwrap(env, argv[0]), p);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}
static napi_value EndGrab(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    int rc = tetsim_end_grab(unwrap(env, argv[0]));
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// skin(handle, visVerts: Float32Array, triIds: Int32Array|null, outPos: Float32Array, outNormals: Float32Array|null)
static napi_value Skin(napi_env env, napi_callback_info info) {
    size_t argc = 5;
    napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    size_t nv = 0, nt = 0, no = 0, nn = 0;
    float *vis = typed<float>(env, argv[1], &nv);
    int32_t *tri = typed<int32_t>(env, argv[2], &nt);
    float *outP = typed<float>(env, argv[3], &no);
    float *outN = typed<float>(env, argv[4], &nn);
    int rc = tetsim_skin(unwrap(env, argv[0]), vis, (int32_t)(nv / 4), tri, (int32_t)(nt / 3), outP, outN);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

static napi_value VolError(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    double v = 0;
    int rc = tetsim_get_vol_error(unwrap(env, argv[0]), &v);
    if (rc != TETSIM_OK) return throw_tetsim(env, rc);
    napi_value out;
    NAPI_OK(napi_create_double(env, v, &out));
    return out;
}

static napi_value Init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"simulate", nullptr, Simulate, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"step", nullptr, Step, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"readPositions", nullptr, Read3<tetsim_get_positions>, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"readPrevPositions", nullptr, Read3<tetsim_get_prev_positions>, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"readVelocities", nullptr, Read3<tetsim_get_velocities>, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"startGrab", nullptr, StartGrab, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"moveGrabbed", nullptr, MoveGrabbed, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"endGrab", nullptr, EndGrab, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"skin", nullptr, Skin, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"volError", nullptr, VolError, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
