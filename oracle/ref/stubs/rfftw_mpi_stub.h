/* Stand-in for FFTW-2's MPI real 3-D transforms (drfftw_mpi.h / srfftw_mpi.h / rfftw_mpi.h are not installed).
 * TEST INFRASTRUCTURE ONLY.  Single rank: the "slab" is the whole mesh.  Data layout and conventions are those of
 * FFTW 2.1.5 (rfftwnd_mpi, in-place): real data [x][y][2*(nz/2+1)] (z padded); after a forward transform with
 * FFTW_TRANSPOSED_ORDER the complex data are [y][x][nz/2+1]; the inverse takes that layout back to the padded real
 * array.  Forward sign exp(-2 pi i jk/n), no normalisation in either direction. */
#ifndef G2_STUB_RFFTW_MPI_H
#define G2_STUB_RFFTW_MPI_H
#include <mpi.h>
#include "fftw_stub.h"
typedef struct g2_rfftwnd_mpi_plan_s { int nx, ny, nz, dir; } *rfftwnd_mpi_plan;
typedef enum { FFTW_NORMAL_ORDER, FFTW_TRANSPOSED_ORDER } fftwnd_mpi_output_order;
#define FFTW_REAL_TO_COMPLEX FFTW_FORWARD
#define FFTW_COMPLEX_TO_REAL FFTW_BACKWARD
#define FFTW_IN_PLACE 8
rfftwnd_mpi_plan rfftw3d_mpi_create_plan(MPI_Comm comm, int nx, int ny, int nz, int dir, int flags);
void rfftwnd_mpi_destroy_plan(rfftwnd_mpi_plan p);
void rfftwnd_mpi_local_sizes(rfftwnd_mpi_plan p, int *local_nx, int *local_x_start, int *local_ny_after_transpose,
			     int *local_y_start_after_transpose, int *total_local_size);
void rfftwnd_mpi(rfftwnd_mpi_plan p, int n_fields, fftw_real *local_data, fftw_real *work, fftwnd_mpi_output_order order);
#endif
