// Functional mock of the subset of MATLAB's MEX / mx API the gateways of matlab/mex use (declarations: mex.h beside this
// file).  Test infrastructure only: arrays are plain heap structs, mexErrMsgIdAndTxt throws, and mock_mex_call() runs a
// gateway's mexFunction and hands the error identifier / message back to the caller (tests drive it with ctypes).
#include "mex.h"
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

struct mxArray_tag {
    mxClassID cls = mxDOUBLE_CLASS;
    bool complex = false, isStruct = false;
    std::vector<mwSize> dims;
    std::vector<unsigned char> data;
    std::vector<std::string> fieldNames;
    std::vector<mxArray*> fields;
};

struct MexError : std::runtime_error {
    std::string id;
    MexError(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};

static size_t elem_size(mxClassID c) {
    switch (c) {
        case mxDOUBLE_CLASS: case mxUINT64_CLASS: return 8;
        case mxSINGLE_CLASS: case mxINT32_CLASS: return 4;
        default: return 1;
    }
}
static size_t numel(const mxArray* a) {
    size_t n = 1;
    for (mwSize d : a->dims) n *= d;
    return n;
}
static mxArray* make(mxClassID cls, bool cplx, mwSize nd, const mwSize* dims) {
    mxArray* a = new mxArray_tag();
    a->cls = cls;
    a->complex = cplx;
    a->dims.assign(dims, dims + nd);
    while (a->dims.size() < 2) a->dims.push_back(1);
    a->data.assign(numel(a) * elem_size(cls) * (cplx ? 2 : 1), 0);
    return a;
}

extern "C" {

bool mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
bool mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
bool mxIsComplex(const mxArray* a) { return a->complex; }
bool mxIsEmpty(const mxArray* a) { return numel(a) == 0; }
bool mxIsStruct(const mxArray* a) { return a->isStruct; }
bool mxIsCell(const mxArray*) { return false; }
size_t mxGetM(const mxArray* a) { return a->dims[0]; }
size_t mxGetN(const mxArray* a) { size_t n = 1; for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i]; return n; }
size_t mxGetNumberOfElements(const mxArray* a) { return numel(a); }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
double mxGetScalar(const mxArray* a) {
    if (numel(a) == 0) throw MexError("mock:mxGetScalar", "empty array");
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *(const double*)a->data.data();
        case mxSINGLE_CLASS: return *(const float*)a->data.data();
        case mxINT32_CLASS: return *(const int32_t*)a->data.data();
        case mxUINT64_CLASS: return (double)*(const uint64_t*)a->data.data();
        default: return a->data[0];
    }
}
double* mxGetDoubles(const mxArray* a) { return (double*)a->data.data(); }
float* mxGetSingles(const mxArray* a) { return (float*)a->data.data(); }
int32_t* mxGetInt32s(const mxArray* a) { return (int32_t*)a->data.data(); }
uint8_t* mxGetUint8s(const mxArray* a) { return (uint8_t*)a->data.data(); }
mxLogical* mxGetLogicals(const mxArray* a) { return (mxLogical*)a->data.data(); }
mxComplexSingle* mxGetComplexSingles(const mxArray* a) { return (mxComplexSingle*)a->data.data(); }
mxComplexDouble* mxGetComplexDoubles(const mxArray* a) { return (mxComplexDouble*)a->data.data(); }
mxArray* mxGetField(const mxArray* s, mwIndex, const char* name) {
    for (size_t i = 0; i < s->fieldNames.size(); ++i)
        if (s->fieldNames[i] == name) return s->fields[i];
    return nullptr;
}
mxArray* mxGetCell(const mxArray*, mwIndex) { return nullptr; }
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { const mwSize d[2] = {m, n}; return make(mxDOUBLE_CLASS, c == mxCOMPLEX, 2, d); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); *mxGetDoubles(a) = v; return a; }
mxArray* mxCreateLogicalMatrix(mwSize m, mwSize n) { const mwSize d[2] = {m, n}; return make(mxLOGICAL_CLASS, false, 2, d); }
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) { const mwSize d[2] = {m, n}; return make(cls, c == mxCOMPLEX, 2, d); }
mxArray* mxCreateNumericArray(mwSize nd, const mwSize* dims, mxClassID cls, mxComplexity c) { return make(cls, c == mxCOMPLEX, nd, dims); }
mxArray* mxCreateStructMatrix(mwSize, mwSize, int nf, const char** names) {
    mxArray* a = new mxArray_tag();
    a->isStruct = true;
    a->dims = {1, 1};
    for (int i = 0; i < nf; ++i) { a->fieldNames.push_back(names[i]); a->fields.push_back(nullptr); }
    return a;
}
void mxSetField(mxArray* s, mwIndex, const char* name, mxArray* v) {
    for (size_t i = 0; i < s->fieldNames.size(); ++i)
        if (s->fieldNames[i] == name) { s->fields[i] = v; return; }
    s->fieldNames.push_back(name);   // the mock also grows structs (convenient for building inputs from Python)
    s->fields.push_back(v);
}
double mxGetNaN(void) { return std::nan(""); }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw MexError(id, buf);
}
static void (*g_atExit)(void) = nullptr;
int mexAtExit(void (*f)(void)) { g_atExit = f; return 0; }
void mexLock(void) {}

// ---- helpers for the test driver ------------------------------------------------------------------------------------
void* mock_data(mxArray* a) { return a->data.data(); }
size_t mock_nbytes(const mxArray* a) { return a->data.size(); }
int mock_class(const mxArray* a) { return (int)a->cls; }
int mock_field_count(const mxArray* a) { return (int)a->fieldNames.size(); }
const char* mock_field_name(const mxArray* a, int i) { return a->fieldNames[i].c_str(); }
void mock_destroy(mxArray* a) {
    if (!a) return;
    for (mxArray* f : a->fields) mock_destroy(f);
    delete a;
}
void mock_at_exit(void) { if (g_atExit) { g_atExit(); g_atExit = nullptr; } }
// run the gateway; 0 = ok, 1 = MATLAB error (identifier / message copied out), 2 = other C++ exception
int mock_mex_call(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs, char* errId, char* errMsg, int cap) {
    try {
        mexFunction(nlhs, plhs, nrhs, prhs);
        return 0;
    } catch (const MexError& e) {
        snprintf(errId, cap, "%s", e.id.c_str());
        snprintf(errMsg, cap, "%s", e.what());
        return 1;
    } catch (const std::exception& e) {
        snprintf(errId, cap, "c++");
        snprintf(errMsg, cap, "%s", e.what());
        return 2;
    }
}

}  // extern "C"
