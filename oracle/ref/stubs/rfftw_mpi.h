#include "rfftw_mpi_stub.h"
