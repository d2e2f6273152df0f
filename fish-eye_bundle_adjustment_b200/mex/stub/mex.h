/* Minimal stand-in for MATLAB's mex.h: declarations only, enough to type-check feba_mex.c where no
 * MATLAB installation exists (the build container).  NOT used when building the real MEX file. */
#ifndef FEBA_STUB_MEX_H
#define FEBA_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;
double* mxGetPr(const mxArray*);
void* mxGetData(const mxArray*);
double mxGetScalar(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
int mxIsDouble(const mxArray*);
int mxIsClass(const mxArray*, const char*);
int mxIsStruct(const mxArray*);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxSetField(mxArray*, mwSize, const char*, mxArray*);
void* mxMalloc(size_t);
void mxFree(void*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexPrintf(const char*, ...);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#endif
