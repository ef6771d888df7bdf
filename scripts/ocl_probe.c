/* Probe: is there a usable OpenCL platform on this box? dlopen only, hand-declared prototypes (no CL headers here). */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
typedef int (*getPlat_t)(unsigned, void**, unsigned*);
typedef int (*getDev_t)(void*, unsigned long, unsigned, void**, unsigned*);
typedef int (*getInfo_t)(void*, unsigned, size_t, void*, size_t*);
int main(int argc, char** argv)
{
  const char* names[] = { argc > 1 ? argv[1] : "libnvidia-opencl.so.1", "libOpenCL.so.1", "libOpenCL.so" };
  for (int k = 0; k < 3; ++k)
  {
    void* L = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    printf("%s: %s\n", names[k], L ? "loaded" : dlerror());
    if (!L)
      continue;
    getPlat_t gp = (getPlat_t)dlsym(L, "clGetPlatformIDs");
    getDev_t gd = (getDev_t)dlsym(L, "clGetDeviceIDs");
    getInfo_t gi = (getInfo_t)dlsym(L, "clGetDeviceInfo");
    getInfo_t pi = (getInfo_t)dlsym(L, "clGetPlatformInfo");
    printf("  symbols: %p %p %p\n", (void*)gp, (void*)gd, (void*)gi);
    if (!gp)
      continue;
    void* plats[8];
    unsigned np = 0;
    int rc = gp(8, plats, &np);
    printf("  clGetPlatformIDs rc=%d n=%u\n", rc, np);
    for (unsigned p = 0; p < np && p < 8; ++p)
    {
      char buf[256] = { 0 };
      if (pi)
        pi(plats[p], 0x0902 /* CL_PLATFORM_NAME */, sizeof buf, buf, 0);
      printf("  platform %u: %s\n", p, buf);
      void* devs[16];
      unsigned nd = 0;
      rc = gd(plats[p], 0xFFFFFFFFul /* ALL */, 16, devs, &nd);
      printf("   clGetDeviceIDs rc=%d n=%u\n", rc, nd);
      for (unsigned d = 0; d < nd && d < 16; ++d)
      {
        memset(buf, 0, sizeof buf);
        gi(devs[d], 0x102B /* CL_DEVICE_NAME */, sizeof buf, buf, 0);
        printf("   device %u: %s\n", d, buf);
      }
    }
  }
  return 0;
}
