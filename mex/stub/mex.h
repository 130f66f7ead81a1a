/* Minimal stand-in for MATLAB's mex.h / matrix.h: ONLY for compile-checking mex/desc_b200_mex.c in
 * an image without MATLAB (tests/test_abi.py).  Declarations follow the documented MEX C API. */
#ifndef DESC_B200_STUB_MEX_H
#define DESC_B200_STUB_MEX_H
#include <stdbool.h>
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsNumeric(const mxArray*);
bool mxIsLogicalScalarTrue(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
double* mxGetPr(const mxArray*);
double mxGetScalar(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#endif
