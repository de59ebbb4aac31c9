// oracle/ref_dmap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference's displacement-map -> normal-map utility *in place* (REF_DMAP_SRC = utils/dmap2nmap.cpp under
// /root/reference; its main() is renamed away with -Dmain=...) and exposes dmap2nmap() on raw planar buffers
// (CImg storage is planar x + y*W + c*W*H, utils/CImg.h:10146-10149).
#include <cstring>
#include REF_DMAP_SRC

extern "C" __attribute__((visibility("default")))
void ref_dmap2nmap(const uint8_t *dmap, int w, int h, float scale, uint8_t *nmap_planar_rgb)
{
	CImg<uint8_t> d(dmap, w, h, 1, 1);
	CImg<uint8_t> n;
	dmap2nmap(d, n, scale);
	memcpy(nmap_planar_rgb, n.data(), (size_t)w * h * 3);
}
