// forwarding header: the reference includes "miscmaths/bfmatrix.h" from FSL ([EXT], absent here); see ../fsl_ext.h
#pragma once
#include "../fsl_ext.h"
