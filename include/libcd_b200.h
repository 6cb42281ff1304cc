/* libcd_b200.h -- libcd-compatible entry points of the B200 SDF build.
 *
 * libcd_b200.so exports, under the reference's own names and with the reference's own
 * struct layout and ownership rules, the grid functions that sit on computedistancefield's
 * path ("lower boundary"):
 *
 *   cd_grid_double_bin_sdf    src/libcd/grid.h:93,  grid.c:637-687
 *   cd_grid_double_dt_sqeuc   src/libcd/grid.h:84,  grid.c:462-569
 *   cd_grid_double_sedt       src/libcd/grid.h:86   (legacy name of the same transform)
 *
 * and one fused replacement for the flood-fill block of the module, whose libcd entry point
 * takes a host callback per cell and so cannot be kept as it is:
 *
 *   cd_grid_b200_flood_relabel   = cd_grid_flood_fill(g, start, 0, replace_1_to_0, 0)
 *                                  + the 1.0 -> HUGE_VAL sweep   (grid_flood.c:30-111,
 *                                  orcdchomp_mod.cpp:143-151, 536-548)
 *
 * Grids are taken from and returned to HOST memory exactly as libcd does: the output grid
 * is a fresh malloc'ed `struct cd_grid` (struct, sizes, lengths and data each their own
 * malloc block, as grid.c:61-132) which the caller releases with libcd's cd_grid_destroy
 * (grid.c:134-143).  Inputs are not modified and not freed.  All work is done on the GPU;
 * without one every call fails (there is no CPU path behind these symbols).
 *
 * Return codes follow libcd: 0 ok, -1 allocation failure, -2 wrong cell type
 * (cell_size != sizeof(double)); additionally -2 for n != 3 (only 3-D grids are
 * accelerated) and -3 for a CUDA / no-device failure; cd_grid_b200_last_error() says which.
 *
 * To use it from the reference: drop the three functions from libcd's grid.c (or link this
 * library first) and link libcd_b200.so + liborcdchomp_b200.so; see INTEGRATION.md section 5.
 */
#ifndef LIBCD_B200_H
#define LIBCD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layout of src/libcd/grid.h:29-41; define LIBCD_B200_NO_STRUCT when libcd's grid.h is
 * included as well */
#ifndef LIBCD_B200_NO_STRUCT
struct cd_grid
{
   int n;            /* dimensionality */
   int *sizes;       /* [n] cells per axis */
   size_t ncells;
   int cell_size;    /* bytes per cell */
   char *data;       /* C order: index = (x * NY + y) * NZ + z */
   double *lengths;  /* [n] side lengths */
};
#endif

int cd_grid_double_bin_sdf(struct cd_grid **gp_dt, struct cd_grid *g_emp);
int cd_grid_double_dt_sqeuc(struct cd_grid **gp_dt, struct cd_grid *g_func);
int cd_grid_double_sedt(struct cd_grid **gp_dt, struct cd_grid *g_func);

/* in place on g (doubles): 6-connected fill 1.0 -> 0.0 from index_start, then every
 * remaining 1.0 -> HUGE_VAL */
int cd_grid_b200_flood_relabel(struct cd_grid *g, size_t index_start);

/* GPU used by the calls above (default: $OCB_DEVICE or 0); takes effect on the next call */
int cd_grid_b200_set_device(int device);
const char *cd_grid_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
