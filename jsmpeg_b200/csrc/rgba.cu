// rgba.cu -- optional epilogue: planar Y/Cr/Cb (coded size) -> RGBA8888 (display size).
//
// Integer-exact restatement of the reference's Canvas2D renderer (src/canvas2d.js:53-122), which
// is what a browser without WebGL displays.  Mind the naming: the decoder calls
// render(Y, Cr, Cb) and the renderer names its parameters (y, cb, cr), so the renderer's `ccb` is
// a Cr sample and its `ccr` a Cb sample:
//     r = Cr + (Cr*103 >> 8) - 179
//     g = (Cb*88 >> 8) - 44 + (Cr*183 >> 8) - 91
//     b = Cb + (Cb*198 >> 8) - 227
//     R = clamp(Y + r), G = clamp(Y - g), B = clamp(Y + b), A = 255
// One chroma pair serves a 2x2 luma quad; only (width>>1) x (height>>1) quads are written, the
// rest of the image keeps the 255 fill of CanvasRenderer.resize (src/canvas2d.js:24-29).
// It runs right after stage 2 on the same stream, so the planes it reads are still in L2.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t clamp_u8(int v) { return (uint32_t)min(255, max(0, v)); }

__device__ __forceinline__ uint32_t rgba_px(int y, int r, int g, int b) {
	return clamp_u8(y + r) | (clamp_u8(y - g) << 8) | (clamp_u8(y + b) << 16) | 0xff000000u;
}

__global__ void rgba_kernel(const ReconTask *__restrict__ tasks) {
	const ReconTask &t = tasks[blockIdx.z];
	if (!t.rgba) return;
	const int qx = blockIdx.x * blockDim.x + threadIdx.x;  // quad column
	const int qy = blockIdx.y * blockDim.y + threadIdx.y;  // quad row
	const int qcols = (t.width + 1) >> 1, qrows = (t.height + 1) >> 1;
	if (qx >= qcols || qy >= qrows) return;
	uint32_t *out = reinterpret_cast<uint32_t *>(t.rgba);
	const int x = qx * 2, y = qy * 2;
	if (qx >= (t.width >> 1) || qy >= (t.height >> 1)) {  // odd edge: untouched by the reference
		for (int dy = 0; dy < 2; dy++)
			for (int dx = 0; dx < 2; dx++)
				if (x + dx < t.width && y + dy < t.height) out[(size_t)(y + dy) * t.width + x + dx] = 0xffffffffu;
		return;
	}
	const int cw = t.coded_width, hw = cw >> 1;
	const int cr = t.cur.cr[qy * hw + qx], cb = t.cur.cb[qy * hw + qx];
	const int r = (cr + ((cr * 103) >> 8)) - 179;
	const int g = ((cb * 88) >> 8) - 44 + ((cr * 183) >> 8) - 91;
	const int b = (cb + ((cb * 198) >> 8)) - 227;
	const uchar2 y01 = *reinterpret_cast<const uchar2 *>(t.cur.y + (size_t)y * cw + x);
	const uchar2 y23 = *reinterpret_cast<const uchar2 *>(t.cur.y + (size_t)(y + 1) * cw + x);
	uint32_t *row0 = out + (size_t)y * t.width + x, *row1 = row0 + t.width;
	row0[0] = rgba_px(y01.x, r, g, b); row0[1] = rgba_px(y01.y, r, g, b);
	row1[0] = rgba_px(y23.x, r, g, b); row1[1] = rgba_px(y23.y, r, g, b);
}

}  // namespace

void launch_rgba(const ReconTask *tasks, int n_tasks, int max_width, int max_height, cudaStream_t stream) {
	if (n_tasks <= 0) return;
	dim3 block(32, 8);
	dim3 grid(((max_width + 1) / 2 + block.x - 1) / block.x, ((max_height + 1) / 2 + block.y - 1) / block.y, n_tasks);
	rgba_kernel<<<grid, block, 0, stream>>>(tasks);
}
