###################################################################################
# Copyright 2021 National Technology & Engineering Solutions of Sandia,           #
# LLC (NTESS). Under the terms of Contract DE-NA0003525 with NTESS, the           #
# U.S. Government retains certain rights in this software.                        #
# If you want to use this code, please refer to the README.rst and LICENSE files. #
###################################################################################

# requires memory_profiler package
# run with  mprof run ./driver.py
# plot with mprof plot [logfile]

try:
    profile
    memRegionsAreEnabled = True
except NameError:
    memRegionsAreEnabled = False
    profile = None
