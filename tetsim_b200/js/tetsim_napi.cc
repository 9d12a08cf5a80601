// tetsim_napi.cc -- N-API shim: exposes include/tetsim_b200.h to Node, one call per entry point the JavaScript wrapper
// classes (softbody.mjs) use.  Build where Node is available:
//   g++ -std=c++17 -shared -fPIC -I<node prefix>/include/node -I../../include tetsim_napi.cc
//       -L.. -ltetsim_b200 -Wl,-rpath,'$ORIGIN/..' -o tetsim_napi.node
// NOT BUILT OR RUN IN THIS REPOSITORY (no Node, no node_api.h, no JS engine in the image); it is type-checked against the
// documented Node-API signatures (node_api_min.h, -DTETSIM_NAPI_TYPECHECK) by tests/test_capi_symbols.py.
// Every typed array coming from JavaScript is checked for element type and length against the handle's mesh before the
// library sees the pointer: a short or wrongly typed array is a thrown TypeError, never a heap overrun.
#ifdef TETSIM_NAPI_TYPECHECK
#include "node_api_min.h"
#else
#include <node_api.h>
#endif

#include <cstring>
#include <string>

#include "tetsim_b200.h"

#define NAPI_OK(call)                                                                  \
    do {                                                                               \
        if ((call) != napi_ok) { napi_throw_error(env, nullptr, "N-API failure: " #call); return nullptr; } \
    } while (0)

static napi_value throw_msg(napi_env env, const std::string &msg) {
    napi_throw_error(env, nullptr, msg.c_str());
    return nullptr;
}
static napi_value throw_tetsim(napi_env env, int rc) {
    // the reference's "error string or null" convention (MultiTargetGPUComputationRenderer.js:178-190) surfaces as a thrown
    // Error carrying the library's message
    return throw_msg(env, std::string("tetsim error ") + std::to_string(rc) + ": " + tetsim_last_error());
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

// A typed array of exactly the expected element type and at least `minLen` elements; `nullable` admits null / undefined.
// Returns false (with a pending exception) on a mismatch.
template <class T>
static bool typed(napi_env env, napi_value v, napi_typedarray_type want, size_t minLen, bool nullable, const char *what, T **out, size_t *len) {
    *out = nullptr;
    *len = 0;
    napi_valuetype vt;
    if (napi_typeof(env, v, &vt) != napi_ok) { throw_msg(env, std::string(what) + ": cannot inspect argument"); return false; }
    if (vt == napi_null || vt == napi_undefined) {
        if (nullable) return true;
        throw_msg(env, std::string(what) + " must not be null");
        return false;
    }
    bool is = false;
    if (napi_is_typedarray(env, v, &is) != napi_ok || !is) { throw_msg(env, std::string(what) + " must be a typed array"); return false; }
    napi_typedarray_type ty;
    void *data = nullptr;
    napi_value ab;
    size_t off = 0, n = 0;
    if (napi_get_typedarray_info(env, v, &ty, &n, &data, &ab, &off) != napi_ok) { throw_msg(env, std::string(what) + ": napi_get_typedarray_info failed"); return false; }
    if (ty != want) { throw_msg(env, std::string(what) + (want == napi_float32_array ? " must be a Float32Array" : " must be an Int32Array")); return false; }
    if (n < minLen) { throw_msg(env, std::string(what) + " holds " + std::to_string(n) + " elements, " + std::to_string(minLen) + " are needed"); return false; }
    *out = static_cast<T *>(data);
    *len = n;
    return true;
}

static tetsim_t *unwrap(napi_env env, napi_value v) {
    void *p = nullptr;
    if (napi_get_value_external(env, v, &p) != napi_ok || !p) { throw_msg(env, "not a tetsim handle"); return nullptr; }
    return static_cast<tetsim_t *>(p);
}
static bool mesh_of(napi_env env, tetsim_t *h, TetSimInfo *info) {
    int rc = tetsim_get_info(h, info);
    if (rc != TETSIM_OK) { throw_tetsim(env, rc); return false; }
    return true;
}

// create(vertices: Float32Array, tetIds: Int32Array, physicsParams, options{solver,arithmetic,iters,...}) -> handle
static napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    if (argc < 3) return throw_msg(env, "create(vertices, tetIds, physicsParams[, options])");
    size_t nv = 0, nt = 0;
    float *verts;
    int32_t *ids;
    if (!typed(env, argv[0], napi_float32_array, 0, false, "vertices", &verts, &nv)) return nullptr;
    if (!typed(env, argv[1], napi_int32_array, 0, false, "tetIds", &ids, &nt)) return nullptr;
    if (nv % 3 || nt % 4) return throw_msg(env, "vertices must hold 3 floats per particle and tetIds 4 ints per element");
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
    tetsim_t *h = unwrap(env, argv[0]);
    if (!h) return nullptr;
    double dt = 0;
    if (napi_get_value_double(env, argv[1], &dt) != napi_ok) return throw_msg(env, "dt must be a number");
    TetSimParams prm;
    read_params(env, argv[2], &prm);
    int rc = tetsim_simulate(h, dt, &prm);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// step(handle, frameDt, numSubsteps, physicsParams)   == the loop at src/main.js:79-84
static napi_value Step(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    if (!h) return nullptr;
    double frameDt = 0, n = 1;
    if (napi_get_value_double(env, argv[1], &frameDt) != napi_ok || napi_get_value_double(env, argv[2], &n) != napi_ok)
        return throw_msg(env, "step(handle, frameDt, numSubsteps, physicsParams)");
    TetSimParams prm;
    read_params(env, argv[3], &prm);
    int rc = tetsim_step(h, frameDt, (int32_t)n, &prm);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// readPositions(handle, out: Float32Array[3 * numParticles]) / readVelocities / readPrevPositions
template <int (*FN)(tetsim_t *, float *)>
static napi_value Read3(napi_env env, napi_callback_info info) {
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    TetSimInfo mi;
    if (!h || !mesh_of(env, h, &mi)) return nullptr;
    size_t n = 0;
    float *out;
    if (!typed(env, argv[1], napi_float32_array, 3 * (size_t)mi.numVerts, false, "out", &out, &n)) return nullptr;
    int rc = FN(h, out);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// readRest(handle, invRestPose[9M] | null, invRestVolume[M] | null, invMass[N] | null)      src/Softbody.js:15-17
static napi_value ReadRest(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    TetSimInfo mi;
    if (!h || !mesh_of(env, h, &mi)) return nullptr;
    size_t n = 0;
    float *q, *v, *m;
    if (!typed(env, argv[1], napi_float32_array, 9 * (size_t)mi.numTets, true, "invRestPose", &q, &n)) return nullptr;
    if (!typed(env, argv[2], napi_float32_array, (size_t)mi.numTets, true, "invRestVolume", &v, &n)) return nullptr;
    if (!typed(env, argv[3], napi_float32_array, (size_t)mi.numVerts, true, "invMass", &m, &n)) return nullptr;
    int rc = tetsim_get_rest(h, q, v, m);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// readPolarState(handle, elems[12M] | null, quats[4M] | null)                              src/SoftbodyGPU.js:54-55
static napi_value ReadPolarState(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    TetSimInfo mi;
    if (!h || !mesh_of(env, h, &mi)) return nullptr;
    size_t n = 0;
    float *r, *q;
    if (!typed(env, argv[1], napi_float32_array, 12 * (size_t)mi.numTets, true, "elems", &r, &n)) return nullptr;
    if (!typed(env, argv[2], napi_float32_array, 4 * (size_t)mi.numTets, true, "quats", &q, &n)) return nullptr;
    int rc = tetsim_get_polar_state(h, r, q);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// startGrab(handle, x, y, z) -> grabId ; moveGrabbed(handle, x, y, z) ; endGrab(handle)
static bool xyz(napi_env env, napi_value *argv, double p[3]) {
    for (int i = 0; i < 3; i++)
        if (napi_get_value_double(env, argv[1 + i], &p[i]) != napi_ok) { throw_msg(env, "x, y, z must be numbers"); return false; }
    return true;
}
static napi_value StartGrab(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    double p[3];
    if (!h || !xyz(env, argv, p)) return nullptr;
    int32_t id = -1;
    int rc = tetsim_start_grab(h, p, &id);
    if (rc != TETSIM_OK) return throw_tetsim(env, rc);
    napi_value out;
    NAPI_OK(napi_create_int32(env, id, &out));
    return out;
}
static napi_value MoveGrabbed(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    double p[3];
    if (!h || !xyz(env, argv, p)) return nullptr;
    int rc = tetsim_move_grabbed(h, p);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}
static napi_value EndGrab(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    if (!h) return nullptr;
    int rc = tetsim_end_grab(h);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// skin(handle, visVerts: Float32Array[4 Nv], triIds: Int32Array[3 Nt] | null, outPos: Float32Array[3 Nv], outNormals | null)
static napi_value Skin(napi_env env, napi_callback_info info) {
    size_t argc = 5;
    napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    if (!h) return nullptr;
    size_t nv = 0, nt = 0, no = 0, nn = 0;
    float *vis, *outP, *outN;
    int32_t *tri;
    if (!typed(env, argv[1], napi_float32_array, 0, false, "visVerts", &vis, &nv)) return nullptr;
    if (nv % 4) return throw_msg(env, "visVerts must hold (tetNr, b0, b1, b2) per surface vertex");
    if (!typed(env, argv[2], napi_int32_array, 0, true, "visTriIds", &tri, &nt)) return nullptr;
    if (nt % 3) return throw_msg(env, "visTriIds must hold 3 indices per triangle");
    if (!typed(env, argv[3], napi_float32_array, 3 * (nv / 4), false, "outPositions", &outP, &no)) return nullptr;
    if (!typed(env, argv[4], napi_float32_array, 3 * (nv / 4), true, "outNormals", &outN, &nn)) return nullptr;
    int rc = tetsim_skin(h, vis, (int32_t)(nv / 4), tri, (int32_t)(nt / 3), outP, outN);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

// skinGpu(handle, visVerts[4 Nv], restNormals[3 Nv] | null, outPos[3 Nv], outNormals[3 Nv] | null)   src/SoftbodyGPU.js:424-448
static napi_value SkinGpu(napi_env env, napi_callback_info info) {
    size_t argc = 5;
    napi_value argv[5];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    if (!h) return nullptr;
    size_t nv = 0, n = 0;
    float *vis, *rn, *outP, *outN;
    if (!typed(env, argv[1], napi_float32_array, 0, false, "visVerts", &vis, &nv)) return nullptr;
    if (nv % 4) return throw_msg(env, "visVerts must hold (tetNr, b0, b1, b2) per surface vertex");
    if (!typed(env, argv[2], napi_float32_array, 3 * (nv / 4), true, "restNormals", &rn, &n)) return nullptr;
    if (!typed(env, argv[3], napi_float32_array, 3 * (nv / 4), false, "outPositions", &outP, &n)) return nullptr;
    if (!typed(env, argv[4], napi_float32_array, 3 * (nv / 4), true, "outNormals", &outN, &n)) return nullptr;
    int rc = tetsim_skin_gpu(h, vis, (int32_t)(nv / 4), rn, outP, outN);
    return rc == TETSIM_OK ? nullptr : throw_tetsim(env, rc);
}

static napi_value VolError(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tetsim_t *h = unwrap(env, argv[0]);
    if (!h) return nullptr;
    double v = 0;
    int rc = tetsim_get_vol_error(h, &v);
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
        {"readRest", nullptr, ReadRest, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"readPolarState", nullptr, ReadPolarState, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"startGrab", nullptr, StartGrab, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"moveGrabbed", nullptr, MoveGrabbed, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"endGrab", nullptr, EndGrab, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"skin", nullptr, Skin, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"skinGpu", nullptr, SkinGpu, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"volError", nullptr, VolError, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
