// wl_gpu_hook.h -- the reference-side binding of scgpu_wl_order / scgpu_wl_mesh (include/scgpu.h): Mesh::meshInit on the device.
//
// WangLandau::runPress and runSwitch recompute the membrane-hole order parameter (wlm 2) from scratch after every volume or type-switch
// move: holeXYPlane(wli) -> Mesh::meshInit = meshFill over all particles + findHoles (scOOP/mc/wanglandau.h:313-317, scOOP/mc/mesh.cpp:11-185).
// With the GPU calculator (totalegpu.h) the committed configuration already lives on the device, so a maintainer changes ONE call:
//
//   scOOP/mc/wanglandau.h, after  #include "mesh.h":      #include "wl_gpu_hook.h"
//   scOOP/mc/wanglandau.h:316  mesh.meshInit(wl_meshsize, ...)  ->  scgpu_mesh_init(&mesh, wl_meshsize, ...)      (same arguments)
//
// The device fills the mesh, finds the largest hole and hands back Mesh::data exactly as findHoles leaves it (hole numbers included),
// because the incremental updates of the next single-particle and chain moves (Mesh::addPart / removePart,
// WangLandau::meshOrderMoveMolecule, scOOP/mc/wanglandau.cpp:7-52) -- which stay host code -- go on from that array.
// oracle/Makefile (target `scgpu_wl_ref`) applies these two edits with sed to a scratch copy and builds oracle/_ref/SC_scgpu_wl;
// tests/test_gpu_dropin.py runs Tests/test_mempore (wlm 2, NPT) through it: config.last and wl-new.dat byte-identical to the unmodified reference.
//
// No include guard around the whole file on purpose: totalegpu.h includes it for the context accessor (Mesh may not be declared yet),
// wanglandau.h includes it again after mesh.h for the mesh part.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scgpu.h"

#ifndef SCGPU_DROPIN_CTX
#define SCGPU_DROPIN_CTX
// the context of the process's GPU calculator (set by TotalEGpu's constructor); one replica per process, as in the reference
inline scgpu_ctx*& scgpu_dropin_ctx() { static scgpu_ctx* c = NULL; return c; }
inline long& scgpu_mesh_init_calls() { static long n = 0; return n; }
#endif

#if defined(MESH_H) && !defined(SCGPU_MESH_HOOK)
#define SCGPU_MESH_HOOK
// Mesh::meshInit(meshsize, npart, wlmtype, box, pvec) (scOOP/mc/mesh.cpp:11-35) with the fill and the hole search on the device.
// The particles are the calculator's committed configuration (every caller of holeXYPlane(wli) has just asked it for allToAll());
// only the box -- which the volume move has already changed on the host -- is sent.
inline int scgpu_mesh_init(Mesh* mesh, double meshsize, long npart, int wlmtype, Vector box, std::vector<Particle>* pvec) {
    (void)npart; (void)pvec;
    scgpu_ctx* ctx = scgpu_dropin_ctx();
    if (!ctx) { fprintf(stderr, "scgpu_mesh_init: no GPU calculator in this process\n"); exit(1); }
    double b[3] = {box.x, box.y, box.z};
    scgpu_wlorder wo;
    memset(&wo, 0, sizeof wo);
    wo.wlm[0] = 2; wo.wlmtype = wlmtype; wo.dorder[0] = 1.0; wo.meshsize = meshsize;
    if (scgpu_set_box(ctx, b) || scgpu_wl_order(ctx, &wo)) { fprintf(stderr, "scgpu_mesh_init: %s\n", scgpu_last_error()); exit(1); }
    const int len = wo.mesh_dim[0] * wo.mesh_dim[1];
    mesh->dim[0] = wo.mesh_dim[0];                    // mesh.cpp:15-21: new dimensions, fresh arrays
    mesh->dim[1] = wo.mesh_dim[1];
    if (mesh->data != NULL) free(mesh->data);
    if (mesh->tmp != NULL) free(mesh->tmp);
    mesh->data = (int*)malloc(sizeof(int) * len);
    mesh->tmp = (int*)malloc(sizeof(int) * (len + 1));
    if (scgpu_wl_mesh(ctx, mesh->data, len)) { fprintf(stderr, "scgpu_mesh_init: %s\n", scgpu_last_error()); exit(1); }
    if (scgpu_mesh_init_calls()++ == 0) fprintf(stderr, "scgpu_mesh_init: Mesh::meshInit runs on the device\n");
    return (int)wo.raw[0];
}
#endif
