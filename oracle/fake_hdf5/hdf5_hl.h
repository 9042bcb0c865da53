// stub of <hdf5_hl.h>: see hdf5.h next to it
#pragma once
