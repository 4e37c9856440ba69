#include "fftw_stub.h"
