/* xm_oracle.h -- TEST INFRASTRUCTURE ONLY (see xm_oracle.c). */
#ifndef XM_ORACLE_H
#define XM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* states and bins share one numbering: the reference's output argument order
 * (xm.py:291-297) */
enum { XMO_PS = 0, XMO_SS = 1, XMO_PM = 2, XMO_SM = 3, XMO_UA = 4, XMO_UR = 5 };
enum { XMO_MODE_SE = 0, XMO_MODE_PE_LIBERAL = 1, XMO_MODE_PE_CONSERVATIVE = 2 };
enum { XMO_SCORE_AS_XS = 0, XMO_SCORE_AS_ZS = 1, XMO_SCORE_CIGAR_NM = 2 };
enum {
    XMO_OK = 0,
    XMO_ERR_ASSERT = 1,      /* AssertionError, xm.py:106 */
    XMO_ERR_VALUE = 2,       /* ValueError, xm.py:190 / float() / int() */
    XMO_ERR_RUNTIME = 3,     /* RuntimeError, xm.py:289 */
    XMO_ERR_UNICODE = 4,     /* UnicodeDecodeError from the 'rt' file layer */
    XMO_ERR_UNSUPPORTED = 5, /* input the oracle does not restate (big ints, non-ASCII digits) */
    XMO_ERR_NOMEM = 6
};

typedef struct {
    int32_t mode;
    int32_t score_src;
    int32_t skip_repeated;
    uint32_t enabled_bins;   /* bit b set = bin b has an output file (None outputs still count, xm.py:330) */
    double min_score;
} xmo_opts;

typedef struct {
    char *data[6];
    size_t len[6];
    uint64_t counts[36];     /* SE: [state]; PE: [fwd*6+rev] */
    uint64_t n_yielded;      /* records the lockstep reader yielded */
    int32_t err;
    uint64_t err_record;
    char errmsg[128];
} xmo_result;

int xmo_classify(const void *prim, size_t prim_len, const void *sec, size_t sec_len,
                 const xmo_opts *opts, xmo_result *res);
void xmo_free(xmo_result *res);
int xmo_mapping_state(double AS1, double XS1, double AS2, double XS2, double min_score);
int xmo_pair_bin(int fwd, int rev, int conservative);
int xmo_line_scores(const void *line, size_t len, int score_src, double *as, double *xs);

#ifdef __cplusplus
}
#endif
#endif
