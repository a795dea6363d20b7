/*
 * rmock.c -- a small stand-in for the part of R's C API that r_shim/baynorm_b200_shim.c uses, so that the shim can be
 * compiled and EXECUTED from the Python tests (no R in this image).  It is test infrastructure: vectors with attributes,
 * S4 objects as attribute bags, external pointers with finalizers, Rf_error as a longjmp back to rmock_call(), the .Call
 * registration table, and a few constructors / accessors for the ctypes harness.  Nothing is ever garbage-collected
 * (objects are leaked until rmock_reset()); finalizers run when the harness asks (rmock_run_finalizers), which is how the
 * tests check that an R error between upload and return does not leak a device matrix.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct attr {
    const char *name;
    SEXP value;
    struct attr *next;
} attr;
struct SEXPREC {
    int type;
    R_xlen_t len;
    void *data; /* double / int / SEXP array; symbol or CHARSXP text; class name */
    attr *attrs;
    void *ext;
    R_CFinalizer_t fin;
};

static struct SEXPREC nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
SEXP R_DimSymbol, R_NamesSymbol;

static SEXP *g_objs = NULL;
static size_t g_nobj = 0, g_cap = 0;
static jmp_buf g_jmp;
static int g_jmp_set = 0;
static char g_err[1024];
static const R_CallMethodDef *g_entries = NULL;
static unsigned long long g_rng = 0x9E3779B97F4A7C15ULL;

static SEXP new_obj(int type, R_xlen_t len, size_t elt)
{
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type;
    s->len = len;
    if (elt) s->data = calloc(len > 0 ? (size_t)len : 1, elt);
    if (g_nobj == g_cap) {
        g_cap = g_cap ? 2 * g_cap : 256;
        g_objs = (SEXP *)realloc(g_objs, g_cap * sizeof(SEXP));
    }
    g_objs[g_nobj++] = s;
    return s;
}

SEXP Rf_install(const char *name)
{
    for (size_t k = 0; k < g_nobj; k++)
        if (g_objs[k]->type == SYMSXP && strcmp((const char *)g_objs[k]->data, name) == 0) return g_objs[k];
    SEXP s = new_obj(SYMSXP, 0, 0);
    s->data = strdup(name);
    return s;
}

static void init_symbols(void)
{
    if (!R_DimSymbol) {
        R_DimSymbol = Rf_install("dim");
        R_NamesSymbol = Rf_install("names");
    }
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n)
{
    init_symbols();
    switch (type) {
    case REALSXP: return new_obj(REALSXP, n, sizeof(double));
    case INTSXP: return new_obj(INTSXP, n, sizeof(int));
    case LGLSXP: return new_obj(LGLSXP, n, sizeof(int));
    case VECSXP: {
        SEXP s = new_obj(VECSXP, n, sizeof(SEXP));
        for (R_xlen_t k = 0; k < n; k++) ((SEXP *)s->data)[k] = R_NilValue;
        return s;
    }
    case STRSXP: return new_obj(STRSXP, n, sizeof(SEXP));
    default: Rf_error("rmock: allocVector of type %u", type);
    }
}
static void set_dim(SEXP s, int n, const int *d)
{
    SEXP dim = Rf_allocVector(INTSXP, n);
    memcpy(dim->data, d, sizeof(int) * (size_t)n);
    Rf_setAttrib(s, R_DimSymbol, dim);
}
SEXP Rf_allocMatrix(unsigned int type, int nr, int nc)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
    const int d[2] = {nr, nc};
    set_dim(s, 2, d);
    return s;
}
SEXP Rf_alloc3DArray(unsigned int type, int a, int b, int c)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)a * b * c);
    const int d[3] = {a, b, c};
    set_dim(s, 3, d);
    return s;
}
SEXP Rf_protect(SEXP s) { return s; }
void Rf_unprotect(int n) { (void)n; }
double *REAL(SEXP s)
{
    if (s->type != REALSXP) Rf_error("rmock: REAL() on a non-double vector (type %d)", s->type);
    return (double *)s->data;
}
int *INTEGER(SEXP s)
{
    if (s->type != INTSXP && s->type != LGLSXP) Rf_error("rmock: INTEGER() on a non-integer vector (type %d)", s->type);
    return (int *)s->data;
}
int *LOGICAL(SEXP s)
{
    if (s->type != LGLSXP) Rf_error("rmock: LOGICAL() on a non-logical vector (type %d)", s->type);
    return (int *)s->data;
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t k) { return ((SEXP *)s->data)[k]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t k, SEXP v) { return ((SEXP *)s->data)[k] = v; }
void SET_STRING_ELT(SEXP s, R_xlen_t k, SEXP v) { ((SEXP *)s->data)[k] = v; }
SEXP Rf_mkChar(const char *t)
{
    SEXP s = new_obj(CHARSXP, (R_xlen_t)strlen(t), 0);
    s->data = strdup(t);
    return s;
}
SEXP Rf_mkString(const char *t)
{
    SEXP s = Rf_allocVector(STRSXP, 1);
    SET_STRING_ELT(s, 0, Rf_mkChar(t));
    return s;
}
SEXP Rf_ScalarReal(double v)
{
    SEXP s = Rf_allocVector(REALSXP, 1);
    REAL(s)[0] = v;
    return s;
}
SEXP Rf_ScalarInteger(int v)
{
    SEXP s = Rf_allocVector(INTSXP, 1);
    INTEGER(s)[0] = v;
    return s;
}
SEXP Rf_getAttrib(SEXP s, SEXP sym)
{
    for (attr *a = s->attrs; a; a = a->next)
        if (strcmp(a->name, (const char *)sym->data) == 0) return a->value;
    return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP sym, SEXP v)
{
    for (attr *a = s->attrs; a; a = a->next)
        if (strcmp(a->name, (const char *)sym->data) == 0) {
            a->value = v;
            return v;
        }
    attr *a = (attr *)calloc(1, sizeof(attr));
    a->name = (const char *)sym->data;
    a->value = v;
    a->next = s->attrs;
    s->attrs = a;
    return v;
}
SEXP R_do_slot(SEXP obj, SEXP sym)
{
    SEXP v = Rf_getAttrib(obj, sym);
    if (v == R_NilValue) Rf_error("no slot of name \"%s\" for this object", (const char *)sym->data);
    return v;
}
SEXP R_do_slot_assign(SEXP obj, SEXP sym, SEXP v)
{
    Rf_setAttrib(obj, sym, v);
    return obj;
}
SEXP R_do_MAKE_CLASS(const char *name)
{
    SEXP s = new_obj(CHARSXP, 0, 0);
    s->data = strdup(name);
    return s;
}
SEXP R_do_new_object(SEXP cls)
{
    SEXP s = new_obj(S4SXP, 0, 0);
    s->data = strdup((const char *)cls->data);
    return s;
}
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (g_jmp_set) longjmp(g_jmp, 1);
    fprintf(stderr, "rmock: Rf_error outside rmock_call: %s\n", g_err);
    abort();
}
int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
int Rf_isReal(SEXP s) { return s->type == REALSXP; }
int Rf_isInteger(SEXP s) { return s->type == INTSXP; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
R_xlen_t Rf_xlength(SEXP s) { return s->len; }
int LENGTH(SEXP s) { return (int)s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
static double elt_as_real(SEXP s, R_xlen_t k)
{
    if (s->type == REALSXP) return ((double *)s->data)[k];
    if (s->type == INTSXP || s->type == LGLSXP) return (double)((int *)s->data)[k];
    Rf_error("rmock: cannot interpret type %d as numeric", s->type);
}
double Rf_asReal(SEXP s)
{
    if (s->len < 1) Rf_error("argument of length 0");
    return elt_as_real(s, 0);
}
int Rf_asInteger(SEXP s) { return (int)Rf_asReal(s); }
int Rf_asLogical(SEXP s) { return Rf_asReal(s) != 0.0; }
SEXP Rf_coerceVector(SEXP s, unsigned int type)
{
    if ((unsigned)s->type == type) return s;
    if (type != REALSXP && type != INTSXP) Rf_error("rmock: coerceVector to type %u", type);
    if (s->type != REALSXP && s->type != INTSXP && s->type != LGLSXP) Rf_error("rmock: cannot coerce type %d", s->type);
    SEXP r = Rf_allocVector(type, s->len);
    for (R_xlen_t k = 0; k < s->len; k++) {
        if (type == REALSXP) ((double *)r->data)[k] = elt_as_real(s, k);
        else ((int *)r->data)[k] = (int)elt_as_real(s, k);
    }
    r->attrs = s->attrs; /* dim, names travel with the data as in R */
    return r;
}
int Rf_nrows(SEXP s)
{
    SEXP d = Rf_getAttrib(s, R_DimSymbol);
    if (d == R_NilValue) Rf_error("object is not a matrix");
    return ((int *)d->data)[0];
}
int Rf_ncols(SEXP s)
{
    SEXP d = Rf_getAttrib(s, R_DimSymbol);
    if (d == R_NilValue || d->len < 2) Rf_error("object is not a matrix");
    return ((int *)d->data)[1];
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag;
    (void)prot;
    SEXP s = new_obj(EXTPTRSXP, 0, 0);
    s->ext = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->ext; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_SetExternalPtrAddr(SEXP s, void *p) { s->ext = p; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit)
{
    (void)onexit;
    s->fin = f;
}
void GetRNGstate(void) {}
void PutRNGstate(void) {}
double unif_rand(void)
{
    g_rng = g_rng * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((double)(g_rng >> 11) + 0.5) / 9007199254740992.0;
}
int R_registerRoutines(DllInfo *dll, const void *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)dll;
    (void)c;
    (void)f;
    (void)e;
    g_entries = call;
    return 1;
}
int R_useDynamicSymbols(DllInfo *dll, Rboolean v)
{
    (void)dll;
    (void)v;
    return 1;
}

/* ---- harness side (called through ctypes) ------------------------------------------------------------------ */
void R_init_bayNorm(DllInfo *dll);
void R_unload_bayNorm(DllInfo *dll);

void rmock_set_seed(unsigned long long s) { g_rng = s * 2654435761ULL + 12345ULL; }
SEXP rmock_nil(void) { return R_NilValue; }
SEXP rmock_real(const double *v, long n)
{
    SEXP s = Rf_allocVector(REALSXP, n);
    if (n) memcpy(s->data, v, sizeof(double) * (size_t)n);
    return s;
}
SEXP rmock_int(const int *v, long n)
{
    SEXP s = Rf_allocVector(INTSXP, n);
    if (n) memcpy(s->data, v, sizeof(int) * (size_t)n);
    return s;
}
SEXP rmock_lgl(int v)
{
    SEXP s = Rf_allocVector(LGLSXP, 1);
    ((int *)s->data)[0] = v;
    return s;
}
void rmock_set_dim2(SEXP s, int nr, int nc)
{
    const int d[2] = {nr, nc};
    set_dim(s, 2, d);
}
SEXP rmock_dgC(int G, int C, SEXP p, SEXP i, SEXP x)
{
    SEXP obj = R_do_new_object(R_do_MAKE_CLASS("dgCMatrix"));
    const int d[2] = {G, C};
    SEXP dim = Rf_allocVector(INTSXP, 2);
    memcpy(dim->data, d, sizeof(d));
    Rf_setAttrib(obj, Rf_install("i"), i);
    Rf_setAttrib(obj, Rf_install("p"), p);
    Rf_setAttrib(obj, Rf_install("x"), x);
    Rf_setAttrib(obj, Rf_install("Dim"), dim);
    return obj;
}
int rmock_type(SEXP s) { return s->type; }
long rmock_len(SEXP s) { return (long)s->len; }
void *rmock_data(SEXP s) { return s->data; }
SEXP rmock_attr(SEXP s, const char *name)
{
    for (attr *a = s->attrs; a; a = a->next)
        if (strcmp(a->name, name) == 0) return a->value;
    return NULL;
}
SEXP rmock_elt(SEXP s, long k) { return ((SEXP *)s->data)[k]; }
const char *rmock_text(SEXP s) { return (const char *)s->data; }
const char *rmock_last_error(void) { return g_err; }

/* .Call(name, args...) through the registered table; NULL + rmock_last_error() when the routine raised an R error */
SEXP rmock_call(const char *name, int nargs, SEXP *a)
{
    if (!g_entries) R_init_bayNorm(NULL);
    const R_CallMethodDef *e = g_entries;
    for (; e->name; e++)
        if (strcmp(e->name, name) == 0) break;
    g_err[0] = 0;
    if (!e->name) {
        snprintf(g_err, sizeof(g_err), "\"%s\" not available for .Call() for package \"bayNorm\"", name);
        return NULL;
    }
    if (e->numArgs != nargs) {
        snprintf(g_err, sizeof(g_err), "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, e->numArgs, name);
        return NULL;
    }
    SEXP r = NULL;
    g_jmp_set = 1;
    if (setjmp(g_jmp) == 0) {
        typedef SEXP (*f1)(SEXP);
        typedef SEXP (*f2)(SEXP, SEXP);
        typedef SEXP (*f3)(SEXP, SEXP, SEXP);
        typedef SEXP (*f4)(SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f5)(SEXP, SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
        typedef SEXP (*f7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
        switch (nargs) {
        case 1: r = ((f1)e->fun)(a[0]); break;
        case 2: r = ((f2)e->fun)(a[0], a[1]); break;
        case 3: r = ((f3)e->fun)(a[0], a[1], a[2]); break;
        case 4: r = ((f4)e->fun)(a[0], a[1], a[2], a[3]); break;
        case 5: r = ((f5)e->fun)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: r = ((f6)e->fun)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: r = ((f7)e->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        default: snprintf(g_err, sizeof(g_err), "rmock: %d arguments", nargs);
        }
    }
    g_jmp_set = 0;
    return r;
}
int rmock_n_routines(void)
{
    if (!g_entries) R_init_bayNorm(NULL);
    int n = 0;
    for (const R_CallMethodDef *e = g_entries; e->name; e++) n++;
    return n;
}
const char *rmock_routine(int k, int *nargs)
{
    if (!g_entries) R_init_bayNorm(NULL);
    *nargs = g_entries[k].numArgs;
    return g_entries[k].name;
}
/* what R's garbage collector would do: run the finalizer of every external pointer; returns how many still held a pointer */
int rmock_run_finalizers(void)
{
    int live = 0;
    for (size_t k = 0; k < g_nobj; k++)
        if (g_objs[k]->type == EXTPTRSXP && g_objs[k]->fin) {
            if (g_objs[k]->ext) live++;
            g_objs[k]->fin(g_objs[k]);
            g_objs[k]->fin = NULL;
        }
    return live;
}
void rmock_unload(void) { R_unload_bayNorm(NULL); }
