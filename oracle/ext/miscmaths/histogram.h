// forwarding header: the reference includes "miscmaths/histogram.h" from FSL ([EXT], absent here); see ../fsl_ext.h
#pragma once
#include "../fsl_ext.h"
