/* Shared helpers of the MEX gateways (R2018a interleaved-complex API: mxComplexSingle == float2).
 * Build (on a machine with MATLAB + CUDA):  mex -R2018a -I../../include isac_fft2d_mex.cpp -L<repo>/.../lib -lisac_b200
 * The gateways only marshal; all arithmetic is behind the C ABI of include/isac_b200.h.  Device buffers for the `_dev`
 * entry points come from isac_dev_malloc / isac_memcpy_* so that no gateway links the CUDA runtime itself. */
#pragma once
#include "isac_b200.h"
#include "mex.h"
#include <cstdint>
#include <string>
#include <vector>

static isac_ctx* g_ctx = nullptr;
static void (*g_plan_cleanup)(void) = nullptr;   /* set by the gateway that caches plans: they go before the context */

static void isac_mex_cleanup(void) {
    if (g_plan_cleanup) g_plan_cleanup();
    if (g_ctx) { isac_destroy(g_ctx); g_ctx = nullptr; }
}

static isac_ctx* isac_mex_ctx(void) {
    if (!g_ctx) {
        int st = isac_create(&g_ctx, 0);
        if (st) mexErrMsgIdAndTxt("isac:create:noDevice", "%s", isac_last_error(nullptr));
        mexAtExit(isac_mex_cleanup);
        mexLock();   /* keep the context (and its device buffers) alive between calls */
    }
    return g_ctx;
}

/* Map a failing status to a MATLAB error so the reference's try/catch (cellSimulation.m:196-202) keeps working. */
static void isac_mex_check(int st, const char* fn) {
    if (st == ISAC_OK) return;
    const std::string id = std::string("isac:") + fn + ":status" + std::to_string(st);
    mexErrMsgIdAndTxt(id.c_str(), "%s", isac_last_error(g_ctx));
}

/* Plans cached per configuration for the life of the MEX file (the context is mexLock'ed): the simulator calls a gateway
 * once per slot / CSI-RS occasion with the same configuration, and creating a plan uploads codebook tables, builds the
 * Gram-pair dictionary and allocates the device arenas.  Keyed by the serialised configuration; at most kMax plans are
 * kept (least recently used goes first). */
template <class Plan>
struct PlanCache {
    struct Entry { std::string key; Plan* plan; };
    static const size_t kMax = 8;
    std::vector<Entry> entries;
    int (*destroy)(Plan*);
    explicit PlanCache(int (*d)(Plan*)) : destroy(d) {}
    Plan* find(const std::string& key) {
        for (size_t i = 0; i < entries.size(); ++i)
            if (entries[i].key == key) {
                Entry hit = entries[i];
                entries.erase(entries.begin() + i);
                entries.push_back(hit);          /* most recently used at the back */
                return hit.plan;
            }
        return nullptr;
    }
    void put(const std::string& key, Plan* plan) {
        if (entries.size() >= kMax) { destroy(entries.front().plan); entries.erase(entries.begin()); }
        entries.push_back(Entry{key, plan});
    }
    void drop(Plan* plan) {                       /* a plan that returned an error is not reused */
        for (size_t i = 0; i < entries.size(); ++i)
            if (entries[i].plan == plan) { destroy(plan); entries.erase(entries.begin() + i); return; }
    }
    void clear() {
        for (Entry& e : entries) destroy(e.plan);
        entries.clear();
    }
};
template <class T>
static void key_add(std::string& k, const T& v) { k.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <class T>
static void key_add(std::string& k, const std::vector<T>& v) {
    const size_t n = v.size();
    key_add(k, n);
    if (n) k.append(reinterpret_cast<const char*>(v.data()), n * sizeof(T));
}

static const mxArray* field(const mxArray* s, const char* name) {
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("isac:mex:missingField", "missing field %s", name);
    return f;
}
static double field_scalar(const mxArray* s, const char* name) { return mxGetScalar(field(s, name)); }

/* device copy of a host array, freed when the object leaves scope */
struct DevBuf {
    void* p = nullptr;
    DevBuf(const void* host, size_t bytes, const char* fn) {
        isac_mex_check(isac_dev_malloc(isac_mex_ctx(), bytes, &p), fn);
        if (host) isac_mex_check(isac_memcpy_h2d(isac_mex_ctx(), p, host, bytes), fn);
    }
    ~DevBuf() { if (p) isac_dev_free(isac_mex_ctx(), p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

static isac_doa_config doa_from_cfg(const mxArray* cfg) {
    isac_doa_config a = {};
    a.isUpa = (int32_t)field_scalar(cfg, "isUpa"); a.nAnts = (int32_t)field_scalar(cfg, "nAnts");
    a.nX = (int32_t)field_scalar(cfg, "nX"); a.nY = (int32_t)field_scalar(cfg, "nY"); a.d = 0.5;   /* music.m:12 */
    a.aGran = field_scalar(cfg, "aGran"); a.aMax = field_scalar(cfg, "aMax");
    a.eGran = field_scalar(cfg, "eGran"); a.eMax = field_scalar(cfg, "eMax");
    return a;
}

static mxArray* row_vector(const double* v, int n) {
    mxArray* m = mxCreateDoubleMatrix(1, (mwSize)n, mxREAL);
    for (int i = 0; i < n; ++i) mxGetDoubles(m)[i] = v[i];
    return m;
}

/* ---- typed accessors ------------------------------------------------------------------------------------------------- */
static void require_csingle(const mxArray* a, const char* fn, const char* what) {
    if (!mxIsSingle(a) || !mxIsComplex(a)) {
        std::string id = std::string("isac:") + fn + ":type";
        mexErrMsgIdAndTxt(id.c_str(), "%s must be complex single (the .m shim converts)", what);
    }
}
static int dim_of(const mxArray* a, int i) {
    return i < (int)mxGetNumberOfDimensions(a) ? (int)mxGetDimensions(a)[i] : 1;
}
static std::vector<double> field_doubles(const mxArray* s, const char* name) {
    const mxArray* f = field(s, name);
    if (!mxIsDouble(f) || mxIsComplex(f)) mexErrMsgIdAndTxt("isac:mex:fieldType", "field %s must be real double", name);
    const double* p = mxGetDoubles(f);
    return std::vector<double>(p, p + mxGetNumberOfElements(f));
}
static std::vector<int32_t> field_int32s(const mxArray* s, const char* name) {
    const mxArray* f = field(s, name);
    const size_t n = mxGetNumberOfElements(f);
    std::vector<int32_t> v(n);
    if (mxIsDouble(f)) { const double* p = mxGetDoubles(f); for (size_t i = 0; i < n; ++i) v[i] = (int32_t)p[i]; }
    else { const int32_t* p = mxGetInt32s(f); for (size_t i = 0; i < n; ++i) v[i] = p[i]; }
    return v;
}
static std::vector<uint8_t> field_bits(const mxArray* s, const char* name) {
    const mxArray* f = mxGetField(s, 0, name);
    std::vector<uint8_t> v;
    if (!f || mxIsEmpty(f)) return v;                 /* empty == no restriction (dlPMISelect.m:744-775) */
    const size_t n = mxGetNumberOfElements(f);
    v.resize(n);
    if (mxIsDouble(f)) { const double* p = mxGetDoubles(f); for (size_t i = 0; i < n; ++i) v[i] = p[i] != 0.0; }
    else { const uint8_t* p = mxGetUint8s(f); for (size_t i = 0; i < n; ++i) v[i] = p[i] != 0; }
    return v;
}
/* complex128 interleaved copy of a (possibly real) double array */
static std::vector<double> complex_doubles(const mxArray* a) {
    const size_t n = mxGetNumberOfElements(a);
    std::vector<double> v(2 * n, 0.0);
    if (mxIsComplex(a)) { const mxComplexDouble* p = mxGetComplexDoubles(a); for (size_t i = 0; i < n; ++i) { v[2*i] = p[i].real; v[2*i+1] = p[i].imag; } }
    else { const double* p = mxGetDoubles(a); for (size_t i = 0; i < n; ++i) v[2*i] = p[i]; }
    return v;
}
static mxArray* double_array(const std::vector<mwSize>& dims, const double* v) {
    mxArray* m = mxCreateNumericArray((mwSize)dims.size(), dims.data(), mxDOUBLE_CLASS, mxREAL);
    const size_t n = mxGetNumberOfElements(m);
    for (size_t i = 0; i < n; ++i) mxGetDoubles(m)[i] = v[i];
    return m;
}
static mxArray* complex_double_array(const std::vector<mwSize>& dims, const double* interleaved) {
    mxArray* m = mxCreateNumericArray((mwSize)dims.size(), dims.data(), mxDOUBLE_CLASS, mxCOMPLEX);
    const size_t n = mxGetNumberOfElements(m);
    mxComplexDouble* p = mxGetComplexDoubles(m);
    for (size_t i = 0; i < n; ++i) { p[i].real = interleaved[2*i]; p[i].imag = interleaved[2*i+1]; }
    return m;
}

/* isac_csi_config + the arrays it points to, from the struct the dlPMISelect / riSelect / cqiSelect shims build out of the
 * validated reportConfig (dlPMISelect.m:511-851) */
struct CsiCfg {
    isac_csi_config c = {};
    std::vector<uint8_t> csr, i2r;
    std::vector<int32_t> reK, reL;
    CsiCfg(const mxArray* cfg, int nRx) {
        c.nPorts = (int32_t)field_scalar(cfg, "nPorts");
        c.N1 = (int32_t)field_scalar(cfg, "N1"); c.N2 = (int32_t)field_scalar(cfg, "N2");
        c.O1 = (int32_t)field_scalar(cfg, "O1"); c.O2 = (int32_t)field_scalar(cfg, "O2");
        c.codebookMode = (int32_t)field_scalar(cfg, "codebookMode");
        c.nSizeBWP = (int32_t)field_scalar(cfg, "nSizeBWP"); c.nStartBWP = (int32_t)field_scalar(cfg, "nStartBWP");
        c.subbandSize = (int32_t)field_scalar(cfg, "subbandSize");
        c.pmiSubband = (int32_t)field_scalar(cfg, "pmiSubband"); c.cqiSubband = (int32_t)field_scalar(cfg, "cqiSubband");
        c.K = (int32_t)field_scalar(cfg, "K"); c.L = (int32_t)field_scalar(cfg, "L");
        c.nRx = nRx;
        csr = field_bits(cfg, "subsetRestriction"); i2r = field_bits(cfg, "i2Restriction");
        c.subsetRestriction = csr.empty() ? nullptr : csr.data();
        c.i2Restriction = i2r.empty() ? nullptr : i2r.data();
        std::vector<uint8_t> rir = field_bits(cfg, "riRestriction");
        for (int i = 0; i < 8; ++i) c.riRestriction[i] = i < (int)rir.size() ? rir[i] : 1;   /* riSelect.m:440-447 */
        reK = field_int32s(cfg, "reK"); reL = field_int32s(cfg, "reL");
        c.nRE = (int32_t)reK.size(); c.reK = reK.data(); c.reL = reL.data();
        const mxArray* np = mxGetField(cfg, 0, "nPanels");   /* optional: Ng of a Type1MultiPanel report (dlPMISelect.m:629-644) */
        c.nPanels = np ? (int32_t)mxGetScalar(np) : 0;
    }
    /* serialised configuration: the key of the gateways' plan caches */
    std::string key() const {
        std::string k;
        const int32_t f[] = {c.nPorts, c.N1, c.N2, c.O1, c.O2, c.codebookMode, c.nSizeBWP, c.nStartBWP, c.subbandSize, c.pmiSubband,
                             c.cqiSubband, c.K, c.L, c.nRx, c.nPanels};
        for (int32_t v : f) key_add(k, v);
        for (int i = 0; i < 8; ++i) key_add(k, c.riRestriction[i]);
        key_add(k, csr); key_add(k, i2r); key_add(k, reK); key_add(k, reL);
        return k;
    }
};
