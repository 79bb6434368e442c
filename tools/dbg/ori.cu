
#include <cstdio>
#include "../../pydegensac_b200/csrc/hgeom.h"
using namespace dg;
__global__ void k(const double* P, int* out) {
  double px1[4], py1[4], px2[4], py2[4], sx1[4], sy1[4], sx2[4], sy2[4];
  for (int t = 0; t < 4; ++t) { px1[t]=P[4*t]; py1[t]=P[4*t+1]; px2[t]=P[4*t+2]; py2[t]=P[4*t+3]; }
  #pragma unroll 1
  for (int t = 0; t < 4; ++t) { sx1[t] = px1[3 - t]; sy1[t] = py1[3 - t]; sx2[t] = px2[3 - t]; sy2[t] = py2[3 - t]; }
  out[0] = oriented_ok_H(sx1, sy1, sx2, sy2);
  double A[4][3], B[4][3], p[3], q[3];
  for (int i=0;i<4;++i){A[i][0]=sx1[i];A[i][1]=sy1[i];A[i][2]=1;B[i][0]=sx2[i];B[i][1]=sy2[i];B[i][2]=1;}
  cross3(p,A[0],A[1]); cross3(q,B[0],B[1]);
  printf("dev t1=%.17g %.17g\n",(p[0]*A[2][0]+p[1]*A[2][1]+p[2]*A[2][2]),(q[0]*B[2][0]+q[1]*B[2][1]+q[2]*B[2][2]));
  printf("dev t2=%.17g %.17g\n",(p[0]*A[3][0]+p[1]*A[3][1]+p[2]*A[3][2]),(q[0]*B[3][0]+q[1]*B[3][1]+q[2]*B[3][2]));
  cross3(p,A[2],A[3]); cross3(q,B[2],B[3]);
  printf("dev t3=%.17g %.17g\n",(p[0]*A[0][0]+p[1]*A[0][1]+p[2]*A[0][2]),(q[0]*B[0][0]+q[1]*B[0][1]+q[2]*B[0][2]));
  printf("dev t4=%.17g %.17g\n",(p[0]*A[1][0]+p[1]*A[1][1]+p[2]*A[1][2]),(q[0]*B[1][0]+q[1]*B[1][1]+q[2]*B[1][2]));
}
int main(){
  double h[16]={453.5498961315202,171.64898408105344,499.7364746270853,138.45452529491428,147.60982826437663,455.2490634977832,603.7887876368243,122.73557650943701,349.2792825723434,500.6893703111864,414.50906704336944,455.4038438420606,322.12402746756226,206.94722114164065,536.7306953458158,105.97746701434467};
  double px1[4],py1[4],px2[4],py2[4],sx1[4],sy1[4],sx2[4],sy2[4];
  for(int t=0;t<4;++t){px1[t]=h[4*t];py1[t]=h[4*t+1];px2[t]=h[4*t+2];py2[t]=h[4*t+3];}
  for(int t=0;t<4;++t){sx1[t]=px1[3-t];sy1[t]=py1[3-t];sx2[t]=px2[3-t];sy2[t]=py2[3-t];}
  printf("host ori=%d\n",(int)oriented_ok_H(sx1,sy1,sx2,sy2));
  double A[4][3], B[4][3], p[3], q[3];
  for (int i=0;i<4;++i){A[i][0]=sx1[i];A[i][1]=sy1[i];A[i][2]=1;B[i][0]=sx2[i];B[i][1]=sy2[i];B[i][2]=1;}
  cross3(p,A[0],A[1]); cross3(q,B[0],B[1]);
  printf("host t1=%.17g %.17g\n",(p[0]*A[2][0]+p[1]*A[2][1]+p[2]*A[2][2]),(q[0]*B[2][0]+q[1]*B[2][1]+q[2]*B[2][2]));
  printf("host t2=%.17g %.17g\n",(p[0]*A[3][0]+p[1]*A[3][1]+p[2]*A[3][2]),(q[0]*B[3][0]+q[1]*B[3][1]+q[2]*B[3][2]));
  double *d; int *o; cudaMalloc(&d,sizeof h); cudaMalloc(&o,4); cudaMemcpy(d,h,sizeof h,cudaMemcpyHostToDevice);
  k<<<1,1>>>(d,o); int r; cudaMemcpy(&r,o,4,cudaMemcpyDeviceToHost); printf("dev ori=%d\n",r); return 0; }
