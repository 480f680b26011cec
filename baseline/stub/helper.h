// stub of CUTLASS examples/common/helper.h (error-check macros the reference file never uses)
