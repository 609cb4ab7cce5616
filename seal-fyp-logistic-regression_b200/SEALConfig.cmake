# SEALConfig.cmake -- makes the reference's `find_package(SEAL)` (CMakeLists.txt:25) resolve to the
# B200 CKKS engine: imported target SEAL::seal = libckks_b200.so + the seal/seal.h shim.
#   cmake -DSEAL_DIR=<repo>/seal-fyp-logistic-regression_b200 <reference> && make
get_filename_component(_ckks_root "${CMAKE_CURRENT_LIST_DIR}" ABSOLUTE)
if(NOT TARGET SEAL::seal)
  add_library(SEAL::seal SHARED IMPORTED)
  set_target_properties(SEAL::seal PROPERTIES
      IMPORTED_LOCATION "${_ckks_root}/libckks_b200.so"
      INTERFACE_INCLUDE_DIRECTORIES "${_ckks_root}/include;${_ckks_root}/../include"
      INTERFACE_COMPILE_FEATURES cxx_std_17)
endif()
set(SEAL_FOUND TRUE)
set(SEAL_VERSION "3.4.5")
