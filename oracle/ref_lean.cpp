// oracle/ref_lean.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference's own LEAN-map utility *in place* (REF_LEAN_SRC is
// utils/nmap2leanmap.cpp or utils/nmap2leanmap_biased.cpp under /root/reference, chosen by the
// Makefile; its main() is renamed away with -Dmain=...) and exposes its nmap2leanmap() /
// check_lean_maps() on raw planar buffers.  CImg's storage is planar x + y*W + c*W*H
// (utils/CImg.h:10146-10149), which is exactly the raw layout used here.
#include REF_LEAN_SRC

extern "C" __attribute__((visibility("default")))
void ref_nmap2leanmap(const uint8_t *nmap_planar_rgb, int w, int h, float base_roughness,
                      float *lean1_planar_rgba, float *lean2_planar_rgba, int run_check)
{
	CImg<uint8_t> nmap(nmap_planar_rgb, w, h, 1, 3);
	CImg<float> l1, l2;
	nmap2leanmap(nmap, l1, l2, base_roughness);
	if (run_check) check_lean_maps(l1, l2);
	memcpy(lean1_planar_rgba, l1.data(), sizeof(float) * (size_t)w * h * 4);
	memcpy(lean2_planar_rgba, l2.data(), sizeof(float) * (size_t)w * h * 4);
}
