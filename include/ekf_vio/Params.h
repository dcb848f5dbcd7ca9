/* Params.h — the tunables the two hot paths read, with the reference's defaults and names
 * (reference include/ekf_vio/Params.h:28-36,46,83-86,103-104 and Params.cpp:17-48).  The other
 * ~35 parameters of the reference are read by nothing on these paths and are not
 * reproduced. */
#ifndef EKFVIO_PARAMS_H_
#define EKFVIO_PARAMS_H_

#define D_INVERSE_IMAGE_SCALE 4                        /* Params.h:28 */
#define D_KILL_PAD 11                                  /* Params.h:33 */
#define D_KLT_MIN_EIGEN 1e-4                           /* Params.h:36 */
#define D_NUM_FEATURES 100                             /* Params.h:46 */
#define D_DEFAULT_POINT_DEPTH 0.5                      /* Params.h:83 */
#define D_DEFAULT_POINT_DEPTH_VARIANCE 100             /* Params.h:84 */
#define D_DEFAULT_POINT_HOMOGENOUS_VARIANCE 0.00001    /* Params.h:86 */
#define D_MAX_PYRAMID_LEVEL 3                          /* Params.h:103 */
#define D_WINDOW_SIZE 21                               /* Params.h:104 */

extern double INVERSE_IMAGE_SCALE;
extern int KILL_PAD;
extern double KLT_MIN_EIGEN;
extern int NUM_FEATURES;
extern double DEFAULT_POINT_DEPTH;
extern double DEFAULT_POINT_DEPTH_VARIANCE;
extern double DEFAULT_POINT_HOMOGENOUS_VARIANCE;
extern int WINDOW_SIZE;
extern int MAX_PYRAMID_LEVEL;

#endif
