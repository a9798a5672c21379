// forwarding header: boost::variant / boost::get stand-in over std::variant; see ../../fsl_ext.h
#pragma once
#include "../../fsl_ext.h"
