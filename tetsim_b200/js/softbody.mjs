// softbody.mjs -- drop-in replacements for the reference's two solver classes
//   SoftBody     (src/Softbody.js:3-298)     and     SoftBodyGPU (src/SoftbodyGPU.js:4-712)
// over the N-API shim (tetsim_napi.cc -> libtetsim_b200.so).  SOURCE ONLY here: this image has no
// JavaScript engine, so these wrappers are not executed in this repository; tetsim_b200/softbody.py is
// their line-for-line Python twin and is what the tests drive.  See INTEGRATION.md for the two-line
// change in src/main.js.
import { createRequire } from 'node:module';
const native = createRequire(import.meta.url)('./tetsim_napi.node');

const SOLVER = { gs_exact: 0, gs_color: 1, jacobi: 2, polar: 3 };
const ARITH = { fast: 0, bitexact: 1 };

class Body {
    // same seven (eight) constructor arguments as the reference, plus an options bag
    constructor(vertices, tetIds, tetEdgeIds, physicsParams, visVerts, visTriIds, visMaterial, world, opts = {}) {
        this.physicsParams = physicsParams;
        this.numParticles = vertices.length / 3;                       // src/Softbody.js:9
        this.numElems = tetIds.length / 4;                             // :10
        this.tetIds = Int32Array.from(tetIds);                         // dragonTetIds is a plain Array (src/Dragon.js:311)
        this.grabId = -1;
        this.grabPos = new Float32Array(3);
        this.visVerts = visVerts;
        this.visTriIds = visTriIds ? Int32Array.from(visTriIds) : null;
        this.numVisVerts = visVerts ? visVerts.length / 4 : 0;
        this._pos = new Float32Array(3 * this.numParticles);
        this._h = native.create(Float32Array.from(vertices), this.tetIds, physicsParams, {
            solver: SOLVER[opts.solver ?? this.constructor.defaultSolver],
            arithmetic: ARITH[opts.arithmetic ?? 'fast'],
            iters: opts.iters ?? 1,
        });
        this.visPositions = new Float32Array(3 * this.numVisVerts);
        this.visNormals = new Float32Array(3 * this.numVisVerts);
        // The caller (or a thin THREE adapter) builds edgeMesh / visMesh from these buffers exactly as
        // src/Softbody.js:36-56 does; the solver itself no longer needs three.js.
    }
    simulate(dt, physicsParams) { native.simulate(this._h, dt, physicsParams ?? this.physicsParams); this._fresh = false; }
    step(physicsParams) {                                              // the loop at src/main.js:79-84 in one call
        const p = physicsParams ?? this.physicsParams;
        native.step(this._h, p.timeScale * p.timeStep, p.numSubsteps, p);
        this._fresh = false;
    }
    get pos() { if (!this._fresh) { native.readPositions(this._h, this._pos); this._fresh = true; } return this._pos; }
    get volError() { return native.volError(this._h); }
    endFrame() { this.updateVisMesh(); }
    updateVisMesh() {
        if (!this.numVisVerts) return;
        const wantN = this.physicsParams.computeNormals !== false && this.visTriIds;
        native.skin(this._h, this.visVerts, wantN ? this.visTriIds : null, this.visPositions, wantN ? this.visNormals : null);
    }
    startGrab(pos) { this.grabId = native.startGrab(this._h, pos.x, pos.y, pos.z); this.grabPos.set([pos.x, pos.y, pos.z]); }
    moveGrabbed(pos) { native.moveGrabbed(this._h, pos.x, pos.y, pos.z); this.grabPos.set([pos.x, pos.y, pos.z]); }
    endGrab() { native.endGrab(this._h); this.grabId = -1; }
}

export class SoftBody extends Body { static defaultSolver = 'gs_exact'; }
export class SoftBodyGPU extends Body {
    static defaultSolver = 'polar';
    simulate(dt, physicsParams) { physicsParams.dt = dt; super.simulate(dt, physicsParams); }   // src/SoftbodyGPU.js:611
    endFrame() { /* src/SoftbodyGPU.js:643-647: the vis mesh is skinned at render time */ }
    readToCPU(_variable, buffer) {                                     // src/SoftbodyGPU.js:649-653 (RGBA stride)
        const p = this.pos;
        for (let i = 0; i < this.numParticles; i++) { buffer[4 * i] = p[3 * i]; buffer[4 * i + 1] = p[3 * i + 1]; buffer[4 * i + 2] = p[3 * i + 2]; }
    }
}
