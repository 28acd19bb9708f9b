/* mock_r.c -- a toy R runtime, TEST INFRASTRUCTURE ONLY: just enough of the R C API (tests/mock_r/Rinternals.h)
 * to load r_shim.c as R would and to drive its .Call entry points from the tests (ctypes).  Objects live in an
 * arena released by mock_reset(); Rf_error longjmps to the dispatcher like R's top level; external-pointer
 * finalizers run at mock_reset() like R's gc at exit. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct mock_sexp {
    int type, n, nrow, ncol;
    void *data;
    R_CFinalizer_t fin;
    struct mock_sexp *next;
};
static struct mock_sexp nil_obj = {NILSXP, 0, 0, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
double R_NaReal;
static struct mock_sexp *arena = NULL;
static struct ralloc { struct ralloc *next; } *rallocs = NULL;
static int protect_depth = 0, protect_imbalance = 0;
static jmp_buf top;
static int top_set = 0;
static char errbuf[1024];
static const R_CallMethodDef *registered = NULL;

static SEXP new_obj(int type, int n, size_t bytes) {
    struct mock_sexp *s = (struct mock_sexp *)calloc(1, sizeof(*s));
    s->type = type; s->n = n; s->nrow = n; s->ncol = 1;
    s->data = bytes ? calloc(1, bytes) : NULL;
    s->next = arena; arena = s;
    return s;
}
static size_t elt_size(int type) {
    switch (type) {
        case LGLSXP: case INTSXP: return sizeof(int);
        case REALSXP: return sizeof(double);
        case VECSXP: case STRSXP: return sizeof(SEXP);
        case RAWSXP: return 1;
        default: return 0;
    }
}
SEXP Rf_protect(SEXP s) { protect_depth++; return s; }
void Rf_unprotect(int n) { protect_depth -= n; if (protect_depth < 0) protect_imbalance = 1; }
SEXP Rf_allocVector(int type, ptrdiff_t n) { return new_obj(type, (int)n, (size_t)n * elt_size(type)); }
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    SEXP s = new_obj(type, nrow * ncol, (size_t)nrow * ncol * elt_size(type));
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
double *REAL(SEXP s) { if (s->type != REALSXP) Rf_error("REAL() on a non-double"); return (double *)s->data; }
int *INTEGER(SEXP s) { if (s->type != INTSXP && s->type != LGLSXP) Rf_error("INTEGER() on a non-integer"); return (int *)s->data; }
int *LOGICAL(SEXP s) { if (s->type != LGLSXP) Rf_error("LOGICAL() on a non-logical"); return (int *)s->data; }
unsigned char *RAW(SEXP s) { if (s->type != RAWSXP) Rf_error("RAW() on a non-raw"); return (unsigned char *)s->data; }
int Rf_asInteger(SEXP s) {
    if (s->n < 1) Rf_error("asInteger of a zero-length object");
    return s->type == REALSXP ? (int)((double *)s->data)[0] : ((int *)s->data)[0];
}
double Rf_asReal(SEXP s) {
    if (s->n < 1) Rf_error("asReal of a zero-length object");
    return s->type == REALSXP ? ((double *)s->data)[0] : (double)((int *)s->data)[0];
}
int Rf_asLogical(SEXP s) { return Rf_asInteger(s) != 0; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); ((int *)s->data)[0] = v; return s; }
SEXP SET_VECTOR_ELT(SEXP v, ptrdiff_t i, SEXP x) { ((SEXP *)v->data)[i] = x; return x; }
SEXP VECTOR_ELT(SEXP v, ptrdiff_t i) { return ((SEXP *)v->data)[i]; }
SEXP STRING_ELT(SEXP v, ptrdiff_t i) { return ((SEXP *)v->data)[i]; }
void SET_STRING_ELT(SEXP v, ptrdiff_t i, SEXP x) { ((SEXP *)v->data)[i] = x; }
SEXP Rf_mkChar(const char *str) {
    SEXP c = new_obj(CHARSXP, (int)strlen(str), strlen(str) + 1);
    strcpy((char *)c->data, str);
    return c;
}
const char *CHAR(SEXP s) { return (const char *)s->data; }
int Rf_length(SEXP s) { return s->n; }
int Rf_nrows(SEXP s) { return s->nrow; }
int Rf_ncols(SEXP s) { return s->ncol; }
int Rf_isReal(SEXP s) { return s->type == REALSXP; }
int Rf_isInteger(SEXP s) { return s->type == INTSXP; }
ptrdiff_t Rf_xlength(SEXP s) { return s->n; }
char *R_alloc(size_t n, int size) {
    struct ralloc *r = (struct ralloc *)calloc(1, sizeof(*r) + n * (size_t)size + 16);
    r->next = rallocs; rallocs = r;
    return (char *)(r + 1);
}
void Rf_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(errbuf, sizeof(errbuf), fmt, ap);
    va_end(ap);
    if (!top_set) { fprintf(stderr, "mock R: error outside .Call: %s\n", errbuf); abort(); }
    longjmp(top, 1);
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = new_obj(EXTPTRSXP, 1, 0); s->data = p; return s; }
void *R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->fin = f; }
int R_registerRoutines(DllInfo *d, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
    (void)d; (void)c; (void)f; (void)e; registered = call; return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *d, Rboolean v) { (void)d; return v; }

/* ---- the side the tests talk to ------------------------------------------------------------------------------- */
void R_init_TADpoleB200(DllInfo *);
void mock_init(void) {
    unsigned long long na = 0x7ff00000000007a2ULL;        /* R's NA_real_: a NaN with payload 1954 */
    memcpy(&R_NaReal, &na, sizeof(na));
    R_init_TADpoleB200(NULL);
}
int mock_registered(const char *name) {                     /* number of arguments, or -1 */
    for (const R_CallMethodDef *m = registered; m && m->name; m++) if (!strcmp(m->name, name)) return m->numArgs;
    return -1;
}
/* .Call(name, args...): NULL and *err filled on error(); checks the registered arity and PROTECT balance */
SEXP mock_dot_call(const char *name, int nargs, SEXP *a, const char **err) {
    typedef SEXP (*f0)(void); typedef SEXP (*f1)(SEXP); typedef SEXP (*f2)(SEXP, SEXP); typedef SEXP (*f3)(SEXP, SEXP, SEXP);
    typedef SEXP (*f4)(SEXP, SEXP, SEXP, SEXP); typedef SEXP (*f5)(SEXP, SEXP, SEXP, SEXP, SEXP);
    typedef SEXP (*f6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
    typedef SEXP (*f8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
    *err = NULL;
    const R_CallMethodDef *m = registered;
    for (; m && m->name; m++) if (!strcmp(m->name, name)) break;
    if (!m || !m->name) { snprintf(errbuf, sizeof(errbuf), "no such routine: %s", name); *err = errbuf; return NULL; }
    if (m->numArgs != nargs) { snprintf(errbuf, sizeof(errbuf), "%s takes %d arguments, got %d", name, m->numArgs, nargs); *err = errbuf; return NULL; }
    protect_depth = 0; protect_imbalance = 0;
    SEXP volatile out = NULL;
    top_set = 1;
    if (setjmp(top) == 0) {
        DL_FUNC f = m->fun;
        switch (nargs) {
            case 0: out = ((f0)f)(); break;
            case 1: out = ((f1)f)(a[0]); break;
            case 2: out = ((f2)f)(a[0], a[1]); break;
            case 3: out = ((f3)f)(a[0], a[1], a[2]); break;
            case 4: out = ((f4)f)(a[0], a[1], a[2], a[3]); break;
            case 5: out = ((f5)f)(a[0], a[1], a[2], a[3], a[4]); break;
            case 6: out = ((f6)f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
            case 8: out = ((f8)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
            default: snprintf(errbuf, sizeof(errbuf), "arity %d not supported by the mock", nargs); *err = errbuf;
        }
        if (!*err && (protect_depth != 0 || protect_imbalance)) {
            snprintf(errbuf, sizeof(errbuf), "%s: PROTECT / UNPROTECT imbalance (%d left)", name, protect_depth);
            *err = errbuf; out = NULL;
        }
    } else {
        *err = errbuf; out = NULL;          /* error(): R unwinds the protect stack itself */
    }
    top_set = 0;
    for (struct ralloc *r = rallocs; r;) { struct ralloc *nx = r->next; free(r); r = nx; }   /* R_alloc lives until .Call returns */
    rallocs = NULL;
    return out;
}
SEXP mock_vector(int type, int n, const void *src) {
    SEXP s = Rf_allocVector(type, n);
    if (src && n) memcpy(s->data, src, (size_t)n * elt_size(type));
    return s;
}
SEXP mock_matrix(int type, int nrow, int ncol, const void *src) {
    SEXP s = Rf_allocMatrix(type, nrow, ncol);
    if (src) memcpy(s->data, src, (size_t)nrow * ncol * elt_size(type));
    return s;
}
SEXP mock_string(const char *str) {
    SEXP c = new_obj(CHARSXP, (int)strlen(str), strlen(str) + 1);
    strcpy((char *)c->data, str);
    SEXP v = Rf_allocVector(STRSXP, 1);
    ((SEXP *)v->data)[0] = c;
    return v;
}
SEXP mock_nil(void) { return R_NilValue; }
int mock_type(SEXP s) { return s->type; }
int mock_len(SEXP s) { return s->n; }
int mock_nrow(SEXP s) { return s->nrow; }
int mock_ncol(SEXP s) { return s->ncol; }
void *mock_data(SEXP s) { return s->data; }
SEXP mock_elt(SEXP s, int i) { return ((SEXP *)s->data)[i]; }
/* end of session: finalizers, then everything is released */
void mock_reset(void) {
    for (struct mock_sexp *s = arena; s; s = s->next) if (s->type == EXTPTRSXP && s->fin && s->data) s->fin(s);
    for (struct mock_sexp *s = arena; s;) {
        struct mock_sexp *nx = s->next;
        if (s->type != EXTPTRSXP) free(s->data);
        free(s);
        s = nx;
    }
    arena = NULL;
}
