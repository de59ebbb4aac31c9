// oracle/stubs/OpenEXR/ImfRgbaFile.h -- TEST INFRASTRUCTURE.
// OpenEXR is not installed in this image.  The reference's utils/nmap2leanmap*.cpp hard-define
// cimg_use_openexr, so CImg.h needs these names to exist in order to COMPILE; the numeric kernel
// we pin against (nmap2leanmap(), utils/nmap2leanmap.cpp:18-54) never touches them.  Every stub
// aborts if it is ever executed.
#ifndef ORACLE_STUB_OPENEXR_H
#define ORACLE_STUB_OPENEXR_H
#include <cstdlib>
struct half {
	half() : v(0) {}
	half(float f) : v(f) {}
	operator float() const { return v; }
	float v; // not a real binary16: the stub is never used for data
};
namespace Imath {
struct V2i { int x, y; };
struct Box2i { V2i min, max; };
}
namespace Imf {
struct Rgba { half r, g, b, a; };
enum RgbaChannels { WRITE_Y, WRITE_YA, WRITE_RGB, WRITE_RGBA };
template <typename T> struct Array2D {
	void resizeErase(long, long) { abort(); }
	T *operator[](long) { abort(); return 0; }
};
struct RgbaInputFile {
	RgbaInputFile(const char *) { abort(); }
	Imath::Box2i dataWindow() const { abort(); return Imath::Box2i(); }
	void setFrameBuffer(Rgba *, long, long) { abort(); }
	void readPixels(int, int) { abort(); }
};
struct RgbaOutputFile {
	RgbaOutputFile(const char *, int, int, RgbaChannels) { abort(); }
	void setFrameBuffer(const Rgba *, long, long) { abort(); }
	void writePixels(int) { abort(); }
};
}
#endif
