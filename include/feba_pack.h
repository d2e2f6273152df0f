/*
 * feba_pack.h -- C ABI of the host-side problem build: the reference's text files
 * -> the numeric structure-of-arrays feba_create() takes.
 *
 * Replaces, for large networks, the part of main.m between "files are read" and "data.points
 * exists":
 *   functions/ReadFiles.m:49   readmatrix(..., 'Delimiter',{' ','\t'}, 'ConsecutiveDelimitersRule','join',
 *                              'LeadingDelimitersRule','ignore','OutputType','string','CommentStyle','#')
 *   main.m:196-258             string -> double (str2double), degrees -> radians, missing distortion
 *                              terms -> 0
 *   main.m:260-264             Estimate_AllGCP: TIE = unique(PHO(:,1))
 *   main.m:277-384             per observation: linear strcmp scans over EXT, INT, CNT and TIE
 *                              (O(n_obs * (nImg + nPts + nTie)) string compares; hours at 10M observations)
 * The scans become hash look-ups (first match wins, as the scans `break` on the first hit) and the
 * .pho file is tokenised and resolved by several host threads.  Host code only: nothing here touches
 * the GPU, and nothing here is on the per-iteration path.
 *
 * Error behaviour follows the reference: a status (0 ok / non-zero) and the text main.m would show
 * in its dialog (main.m:293-297, :316-320, :352-356) available from feba_pack_last_error().
 * str2double: text that is not a decimal number (optional sign, digits, '.', exponent, inf, nan)
 * becomes NaN (MATLAB's further forms -- '1,000', '1d3', complex -- are not recognised and give NaN).
 */
#ifndef FEBA_PACK_H
#define FEBA_PACK_H

#include <stddef.h>
#include <stdint.h>

#include "feba.h"

#ifdef __cplusplus
extern "C" {
#endif

enum feba_pack_status {
    FEBA_PACK_OK = 0,
    FEBA_PACK_ERR_IO = 1,        /* a file cannot be read ("Error reading files", main.m:100-104)   */
    FEBA_PACK_ERR_FORMAT = 2,    /* .int without two rows per camera, bad arguments                 */
    FEBA_PACK_ERR_LOOKUP = 3,    /* image / camera / target ID not found (main.m:293,316,352)       */
    FEBA_PACK_ERR_LAYOUT = 4     /* EXT / INT rows do not start with the images / cameras PHO uses
                                    (Buildxhat.m:22-30 takes rows 1..numImg, 1..numCam as slots)    */
};

typedef struct feba_pack feba_pack;

/* Views into the packed problem; valid until feba_pack_free().  Same meaning as the fields of
 * feba_problem (feba.h); n_iop_cols = 3 + NK + 2; tie_pt[t] = CNT row of TIE entry t or -1
 * (Buildxhat.m:110-122). */
typedef struct feba_pack_view {
    int64_t n_obs;
    int32_t n_img, n_cam, n_pts, n_tie, n_iop_cols, reserved;
    const double *obs_x, *obs_y;
    const int32_t *obs_img, *obs_pt;
    const int32_t *img_cam;
    const double *eop0, *iop0, *cam_box, *xyz0;
    const int32_t *pt_tie, *tie_pt;
} feba_pack_view;

/* Read and resolve one data set.  `tie` may be NULL (no tie points) and is ignored when all_gcp != 0
 * (main.m:260-264).  num_radial = Num_Radial_Distortions (columns 4.. of the second .int row,
 * main.m:243-253).  threads <= 0: one per host core (at most 32). */
int feba_pack_read(const char *pho, const char *ext, const char *cnt, const char *intr, const char *tie,
                   int32_t num_radial, int32_t all_gcp, int32_t threads, feba_pack **out);
int feba_pack_get(const feba_pack *p, feba_pack_view *view);
/* Fill a feba_problem (pointers into the pack) ready for feba_create(). */
int feba_pack_problem(const feba_pack *p, const feba_settings *settings, feba_problem *problem);
/* ID strings, '\n'-separated, of table which = 0 targets (CNT rows), 1 images (first n_img EXT rows),
 * 2 cameras (first n_cam INT cameras), 3 TIE entries.  Writes at most cap bytes, returns the size
 * needed in *need (call with cap = 0 first). */
enum { FEBA_PACK_IDS_TARGET = 0, FEBA_PACK_IDS_IMAGE = 1, FEBA_PACK_IDS_CAMERA = 2, FEBA_PACK_IDS_TIE = 3 };
int feba_pack_ids(const feba_pack *p, int32_t which, char *out, size_t cap, size_t *need);
/* Seconds spent in the stages of the last feba_pack_read of this pack: [0] small tables + hash maps,
 * [1] .pho tokenise + resolve, [2] tie bookkeeping. */
int feba_pack_timing(const feba_pack *p, double sec[3]);
const char *feba_pack_last_error(void); /* text of the last failed call on this thread */
void feba_pack_free(feba_pack *p);

#ifdef __cplusplus
}
#endif
#endif /* FEBA_PACK_H */
