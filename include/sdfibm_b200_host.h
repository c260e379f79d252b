/*
 * sdfibm_b200_host.h — C binding of the host façade (sdfibm_b200/host/, class sdfibm::SolidCloud).
 *
 * The façade keeps the reference's C++ surface (src/solidcloud.h:89-123: constructor from a solidDict file and the
 * velocity field, interact / evolve / saveState / fixInternal / saveRestart) and is what an OpenFOAM build links.
 * These entry points expose the same calls over plain C for hosts that are not C++ (the Python test-suite, the
 * stand-alone runner): one call per SolidCloud member main.cpp uses (src/main.cpp:38-39,66,82-83,87,101).
 * Every entry returns 0 on success; sdfibm_host_last_error() gives the message of the last failure on this thread.
 * The compute itself happens in libsdfibm_b200.so (include/sdfibm_b200.h); there is no CPU fallback.
 */
#ifndef SDFIBM_B200_HOST_H
#define SDFIBM_B200_HOST_H

#include "sdfibm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdfibm_host_cloud sdfibm_host_cloud;

const char *sdfibm_host_last_error(void);

/* SolidCloud(dictfile, U, time) (src/solidcloud.cpp:209-274) on a Foam-free mesh: registers U, As, Fs, Ts, Ct on the
 * mesh, takes rho from `rho_fluid` (constant/transportProperties), writes cloud.out / cloud.log under case_dir.
 * U_init[3*n_cells] may be NULL (zero field).  start_time > 0 means restart: initialCorrect() is skipped. */
int sdfibm_host_create(const char *dictfile, const char *case_dir, const sdfibm_mesh_t *mesh, double rho_fluid,
                       double start_time, const double *U_init, sdfibm_host_cloud **out);
int sdfibm_host_destroy(sdfibm_host_cloud *h);

/* cell values of a registered field: "U", "Fs" (3 per cell), "As", "Ts", "Ct" (1 per cell) */
int sdfibm_host_field(sdfibm_host_cloud *h, const char *name, double **data, int64_t *n_values);

int sdfibm_host_is_on_fluid(sdfibm_host_cloud *h, int *on_fluid, int *on_twod);
int sdfibm_host_interact(sdfibm_host_cloud *h, double time, double dt);      /* src/main.cpp:66  */
int sdfibm_host_evolve(sdfibm_host_cloud *h, double time, double dt);        /* src/main.cpp:82  */
int sdfibm_host_save_state(sdfibm_host_cloud *h);                            /* src/main.cpp:83  */
int sdfibm_host_fix_internal(sdfibm_host_cloud *h, double dt);               /* src/main.cpp:87  */
int sdfibm_host_save_restart(sdfibm_host_cloud *h, const char *filename);    /* src/main.cpp:101 */

/* tool_vof (tool_vof/main.cpp, tool_vof/solidcloud.cpp:141-173): VofCloud(dictfile, mesh).writeVOF(field_name) — the volume
 * fraction of the union of the `solids{}` then `planes{}` bodies of a tool_vof-flavoured solidDict; writes
 * <case_dir>/0_<field_name>; alpha[n_cells], total_volume, n_solids, n_planes may be NULL. */
int sdfibm_host_write_vof(const char *dictfile, const char *case_dir, const sdfibm_mesh_t *mesh, const char *field_name,
                          double *alpha, double *total_volume, int *n_solids, int *n_planes);

/* solid states: rigid-body records, total (force, torque) of the last evolve sub-iteration [6N], and the fluid
 * (force, torque) of the last interact [6N] */
int sdfibm_host_n_solids(sdfibm_host_cloud *h, int *n);
int sdfibm_host_get_solids(sdfibm_host_cloud *h, sdfibm_solid_t *out);
int sdfibm_host_get_forces(sdfibm_host_cloud *h, double *force_torque, double *fluid_force_torque);
int sdfibm_host_get_masses(sdfibm_host_cloud *h, double *mass);

/* SolidCloud::calcMeanField with the sampler shape of meta.sampler: mean[3N] (src/solidcloud.cpp:315-359) */
int sdfibm_host_mean_field(sdfibm_host_cloud *h, double *mean);

/* UGrid cell size of the collision step (HEAD: -2 => no pairs) and the process-wide sub-iteration count reset */
int sdfibm_host_set_collision_delta(sdfibm_host_cloud *h, double delta);
int sdfibm_host_reset_subiterations(void);

/* plugin registries: kind = "shape" | "motion" | "forcer"; *found = 1 if `type_name` is registered */
int sdfibm_host_factory_has(const char *kind, const char *type_name, int *found);
/* registers a shape type "TestNoDeviceTag" whose lower() is not implemented (to exercise the hard error) */
int sdfibm_host_register_untagged_shape(void);
/* build the shape `shape_name` of a solidDict: its device record and mass properties
 * props[0..5] = volume, volumeINV, radiusB, moi xx, yy, zz */
int sdfibm_host_shape_record(const char *dictfile, const char *shape_name, sdfibm_shape_t *record, double props[6]);
/* host-side IShape::phi01 / phi at n points for (shape of dictfile, pos, quat) — diagnostics */
int sdfibm_host_shape_eval(const char *dictfile, const char *shape_name, const double pos[3], const double quat[4],
                           const double *points, int64_t n, int32_t *inside, double *phi);

#ifdef __cplusplus
}
#endif
#endif /* SDFIBM_B200_HOST_H */
