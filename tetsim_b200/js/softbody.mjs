// softbody.mjs -- drop-in replacements for the reference's two solver classes
//   SoftBody     (src/Softbody.js:3-298)     and     SoftBodyGPU (src/SoftbodyGPU.js:4-712)
// over the N-API shim (tetsim_napi.cc -> libtetsim_b200.so).  Same constructor arguments, same methods, same fields the
// callers touch: main.js adds .edgeMesh / .visMesh to the scene (src/main.js:67-68), the grabbers raycast them and reach the
// body through .userData (src/Softbody.js:445-451, src/SoftbodyGPU.js:792-806), GPUGrabber calls .updateEdgeMesh() first
// (src/SoftbodyGPU.js:792).  INTEGRATION.md shows the two-line change in src/main.js.
//
// NOT EXECUTED IN THIS REPOSITORY: the build image has no JavaScript engine (no node / deno / bun / browser).  The file is
// kept in lock-step with tetsim_b200/softbody.py, its Python twin, which is what the tests drive; the shim underneath is
// type-checked against the Node-API signatures by tests/test_capi_symbols.py.
import * as THREE from 'three';                       // index.html:36-38 maps 'three' for the browser build; Node resolves the package
import { createRequire } from 'node:module';
const native = createRequire(import.meta.url)('./tetsim_napi.node');

const SOLVER = { gs_exact: 0, gs_color: 1, jacobi: 2, polar: 3 };
const ARITH = { fast: 0, bitexact: 1 };

class Body {
    // the reference's seven (eight) constructor arguments, plus an options bag
    constructor(vertices, tetIds, tetEdgeIds, physicsParams, visVerts, visTriIds, visMaterial, world, opts = {}) {
        this.physicsParams = physicsParams;
        this.numParticles = vertices.length / 3;                       // src/Softbody.js:9
        this.numElems = tetIds.length / 4;                             // :10
        this.tetIds = Int32Array.from(tetIds);                         // dragonTetIds is a plain Array (src/Dragon.js:311)
        this.grabId = -1;                                              // :23
        this.grabPos = new Float32Array(3);                            // :22
        this.world = world;
        this._h = native.create(Float32Array.from(vertices), this.tetIds, physicsParams, {
            solver: SOLVER[opts.solver ?? this.constructor.defaultSolver],
            arithmetic: ARITH[opts.arithmetic ?? 'fast'],
            iters: opts.iters ?? 1,
            clusterSize: opts.clusterSize ?? 256,
            referenceTableBug: opts.referenceTableBug ?? 1,
            device: opts.device ?? -1,
        });
        this._state = { pos: new Float32Array(3 * this.numParticles), prevPos: new Float32Array(3 * this.numParticles),
                        vel: new Float32Array(3 * this.numParticles) };
        this._fresh = { pos: false, prevPos: false, vel: false };
        this._rest = null;

        // visual edge mesh, src/Softbody.js:36-42 / src/SoftbodyGPU.js:415-421: the position attribute wraps the CALLER's
        // `vertices` array, exactly as the reference does (it is overwritten by updateEdgeMesh)
        this.geometry = new THREE.BufferGeometry();
        this.geometry.setAttribute('position', new THREE.BufferAttribute(vertices, 3));
        this.geometry.setIndex(Array.from(tetEdgeIds ?? []));
        this.edgeMesh = new THREE.LineSegments(this.geometry);
        this.edgeMesh.userData = this;                                 // for raycasting
        this.edgeMesh.layers.enable(1);
        this.edgeMesh.visible = true;

        // visual embedded mesh, src/Softbody.js:46-57
        this.visVerts = visVerts ?? new Float32Array(0);
        this.visTriIds = visTriIds ? Int32Array.from(visTriIds) : new Int32Array(0);
        this.numVisVerts = this.visVerts.length / 4;
        this.geometry = new THREE.BufferGeometry();
        this.geometry.setAttribute('position', new THREE.BufferAttribute(new Float32Array(3 * this.numVisVerts), 3));
        this.geometry.setAttribute('normal', new THREE.BufferAttribute(new Float32Array(3 * this.numVisVerts), 3));
        this.geometry.setIndex(Array.from(this.visTriIds));
        this.visMesh = new THREE.Mesh(this.geometry, visMaterial);
        this.visMesh.castShadow = true;
        this.visMesh.userData = this;                                  // for raycasting
        this.visMesh.layers.enable(1);
        this.updateVisMesh();                                          // skinning + vertex normals on the device (:56-57)
    }

    // ---- the hot path ----
    simulate(dt, physicsParams) {                                      // src/Softbody.js:195 / src/SoftbodyGPU.js:610: ONE substep
        native.simulate(this._h, dt, physicsParams ?? this.physicsParams);
        this._stale();
    }
    step(physicsParams) {                                              // the loop at src/main.js:79-84 as one CUDA-graph launch
        const p = physicsParams ?? this.physicsParams;
        native.step(this._h, p.timeScale * p.timeStep, p.numSubsteps, p);
        this._stale();
    }
    _stale() { this._fresh.pos = this._fresh.prevPos = this._fresh.vel = false; }
    _read(name, fn) { if (!this._fresh[name]) { fn(this._h, this._state[name]); this._fresh[name] = true; } return this._state[name]; }

    // ---- readable state, src/Softbody.js:12-20 ----
    get pos() { return this._read('pos', native.readPositions); }
    get prevPos() { return this._read('prevPos', native.readPrevPositions); }
    get vel() { return this._read('vel', native.readVelocities); }
    _restData() {
        if (!this._rest) {
            this._rest = { invRestPose: new Float32Array(9 * this.numElems), invRestVolume: new Float32Array(this.numElems),
                           invMass: new Float32Array(this.numParticles) };
            native.readRest(this._h, this._rest.invRestPose, this._rest.invRestVolume, this._rest.invMass);
        }
        return this._rest;
    }
    get invRestPose() { return this._restData().invRestPose; }
    get invRestVolume() { return this._restData().invRestVolume; }
    get invMass() { return this._restData().invMass; }
    get volError() { return native.volError(this._h); }

    // ---- frame end / render buffers, src/Softbody.js:244-277 ----
    endFrame() { this.updateEdgeMesh(); this.updateVisMesh(); }
    updateEdgeMesh() {                                                 // src/Softbody.js:249-257, src/SoftbodyGPU.js:655-668
        const attr = this.edgeMesh.geometry.attributes.position;
        attr.array.set(this.pos);
        attr.needsUpdate = true;
        this.edgeMesh.geometry.computeBoundingSphere();
    }
    updateVisMesh() {                                                  // src/Softbody.js:259-277; normals replace computeVertexNormals()
        if (!this.numVisVerts) return;
        const g = this.visMesh.geometry;
        const wantN = this.physicsParams.computeNormals !== false && this.visTriIds.length > 0;
        native.skin(this._h, this.visVerts, wantN ? this.visTriIds : null, g.attributes.position.array, wantN ? g.attributes.normal.array : null);
        g.attributes.position.needsUpdate = true;
        if (wantN) g.attributes.normal.needsUpdate = true;
        g.computeBoundingSphere();
    }

    // ---- grab, src/Softbody.js:279-298 ----
    startGrab(pos) { this.grabId = native.startGrab(this._h, pos.x, pos.y, pos.z); this.grabPos.set([pos.x, pos.y, pos.z]); }
    moveGrabbed(pos) { native.moveGrabbed(this._h, pos.x, pos.y, pos.z); this.grabPos.set([pos.x, pos.y, pos.z]); }
    endGrab() { native.endGrab(this._h); this.grabId = -1; }
}

export class SoftBody extends Body { static defaultSolver = 'gs_exact'; }

export class SoftBodyGPU extends Body {
    static defaultSolver = 'polar';
    constructor(vertices, tetIds, tetEdgeIds, physicsParams, visVerts, visTriIds, visMaterial, world, opts = {}) {
        super(vertices, tetIds, tetEdgeIds, physicsParams, visVerts, visTriIds, visMaterial, world, opts);
        // computeVertexNormals() + updateVisMesh() on the rest pose end the reference's constructor (src/SoftbodyGPU.js:484-485):
        // those normals are the `objectNormal` its vertex shader rotates by the tets' quaternions
        this.restNormals = Float32Array.from(this.visMesh.geometry.attributes.normal.array);
    }
    simulate(dt, physicsParams) { physicsParams.dt = dt; super.simulate(dt, physicsParams); }   // src/SoftbodyGPU.js:611
    endFrame() {                                                       // src/SoftbodyGPU.js:643-647
        this.edgeMesh.visible = this.physicsParams.ShowTetMesh;
        this.renderVisMesh();                                          // the patched vertex shader's work (:424-448), done by the library
    }
    renderVisMesh() {
        if (!this.numVisVerts) return;
        const g = this.visMesh.geometry;
        native.skinGpu(this._h, this.visVerts, this.restNormals, g.attributes.position.array, g.attributes.normal.array);
        g.attributes.position.needsUpdate = true;
        g.attributes.normal.needsUpdate = true;
    }
    readToCPU(variable, buffer) {                                      // src/SoftbodyGPU.js:649-653: RGBA-strided floats
        const src = variable === 'vel' ? this.vel : variable === 'prevPos' ? this.prevPos : this.pos;
        for (let i = 0; i < this.numParticles; i++) { buffer[4 * i] = src[3 * i]; buffer[4 * i + 1] = src[3 * i + 1]; buffer[4 * i + 2] = src[3 * i + 2]; }
    }
    get quats() { const q = new Float32Array(4 * this.numElems); native.readPolarState(this._h, null, q); return q; }
    get elems() { const r = new Float32Array(12 * this.numElems); native.readPolarState(this._h, r, null); return r; }
}
