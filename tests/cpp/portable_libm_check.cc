#include "../../quadrotorilqr_b200/csrc/qilqr_portable_libm.h"
#include <cstdio>
#include <cmath>
#include <random>
#include <cstring>
#include <cstdint>
static double ulpdiff(double a, double b){ if (a==b) return 0; double u = std::nextafter(std::fabs(b), INFINITY)-std::fabs(b); return std::fabs(a-b)/u; }
int main(){
  std::mt19937_64 g(1); std::uniform_real_distribution<double> U(-7.0,7.0), U1(-1,1);
  double ms=0,mc=0,ma=0; long ns=0,nc=0,na=0; const int n=4000000;
  for(int i=0;i<n;i++){ double x=U(g); if(i%3==0) x*=1e-3; double s,c; qilqr_plibm::sincos(x,&s,&c);
    long double ls=sinl((long double)x), lc=cosl((long double)x);
    double es=ulpdiff(s,(double)ls), ec=ulpdiff(c,(double)lc); if(es>ms)ms=es; if(ec>mc)mc=ec; ns+= (s!=std::sin(x)); nc+=(c!=std::cos(x));
    double y=std::fabs(U1(g)), w=U1(g); double a=qilqr_plibm::atan2(y,w); long double la=atan2l((long double)y,(long double)w); double ea=ulpdiff(a,(double)la); if(ea>ma)ma=ea; na+=(a!=std::atan2(y,w)); }
  printf("max ulp err vs long double: sin %.2f cos %.2f atan2 %.2f ; differ from glibc: sin %.3f%% cos %.3f%% atan2 %.3f%%\n",ms,mc,ma,100.0*ns/n,100.0*nc/n,100.0*na/n);
  printf("%a %a %a %a\n", qilqr_plibm::atan2(0.0,-1.0), qilqr_plibm::atan2(1.0,0.0), qilqr_plibm::atan2(-0.0,1.0), qilqr_plibm::atan2(1e-300,1.0));
}
