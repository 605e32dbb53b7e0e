/* Minimal stand-in for MATLAB's mex.h (R2018a interleaved-complex API).
 * MATLAB is not available in the build image, so the gateways under matlab/mex cannot be compiled into MEX files here.
 * tests/test_mex_sources_cpu.py type-checks each of them against this header and include/isac_b200.h, and
 * tests/test_mex_mock_*.py build them together with mex_mock.cpp (a small functional implementation of these functions on a
 * plain C struct) into shared objects, so that the gateways are linked against libisac_b200.so and EXECUTED end to end --
 * on the CPU up to the "no CUDA device" error, on a B200 against the Python mirror's results. */
#pragma once
#include <cstddef>
#include <cstdint>

typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef bool mxLogical;
typedef struct { float real, imag; } mxComplexSingle;
typedef struct { double real, imag; } mxComplexDouble;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxLOGICAL_CLASS = 3, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7, mxUINT8_CLASS = 9, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;

#ifdef __cplusplus
extern "C" {
#endif
bool mxIsSingle(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsCell(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
double mxGetScalar(const mxArray*);
double* mxGetDoubles(const mxArray*);
float* mxGetSingles(const mxArray*);
int32_t* mxGetInt32s(const mxArray*);
uint8_t* mxGetUint8s(const mxArray*);
mxLogical* mxGetLogicals(const mxArray*);
mxComplexSingle* mxGetComplexSingles(const mxArray*);
mxComplexDouble* mxGetComplexDoubles(const mxArray*);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
mxArray* mxGetCell(const mxArray*, mwIndex);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateLogicalMatrix(mwSize, mwSize);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
double mxGetNaN(void);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexAtExit(void (*)(void));
void mexLock(void);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
