#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <chrono>
#include <thread>
extern "C" size_t tc_pack2bit_range(const uint8_t*, uint64_t, uint64_t, uint64_t, uint8_t*, uint32_t*, uint8_t*, size_t);
int main(int argc, char**argv){
  size_t n = 1ull<<30; int T = argc>1?atoi(argv[1]):8;
  std::vector<uint8_t> s(n), out(n/4+8);
  const char *al="ACGTacgt"; uint64_t x=88172645463325252ull;
  for(size_t i=0;i<n;i++){ x^=x<<13;x^=x>>7;x^=x<<17; s[i]=al[x&7]; }
  s[12345]='N'; s[999999]='n'; s[5]='R';
  // correctness vs scalar decode
  std::vector<uint32_t> ep(16); std::vector<uint8_t> eb(16);
  size_t ne = tc_pack2bit_range(s.data(),0,1<<20,0,out.data(),ep.data(),eb.data(),16);
  printf("exceptions %zu: %u %u %u\n", ne, ep[0],ep[1],ep[2]);
  const char dec[4]={'A','C','T','G'}; size_t bad=0;
  for(size_t i=0;i<(1<<20);i++){ uint8_t c=(out[i/4]>>(2*(i&3)))&3; uint8_t want=s[i]&0xDF; if(dec[c]!=want && i!=12345&&i!=999999&&i!=5) bad++; }
  printf("mismatches %zu\n", bad);
  for(int rep=0;rep<3;rep++){
  auto t0=std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for(int t=0;t<T;t++) th.emplace_back([&,t]{ size_t a=(n/T*t)&~3ull, b=t==T-1?n:(n/T*(t+1))&~3ull; uint32_t p[16]; uint8_t q[16]; tc_pack2bit_range(s.data(),a,b,0,out.data()+a/4,p,q,16); });
  for(auto&t:th)t.join();
  double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  printf("T=%d: %.2f GB/s of ASCII\n", T, n/dt/1e9);}
}
