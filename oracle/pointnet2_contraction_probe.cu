// TEST INFRASTRUCTURE -- probe, not product. The loop body of the published pointnet2_ops FPS kernel
// (erikwijmans/Pointnet2_PyTorch, pointnet2_ops_lib 3.0.0, _ext-src/src/sampling_gpu.cu; an un-vendored dependency of the
// reference, models/point_encoder.py:3,12), restated only to read off how nvcc contracts its two sums:
//   nvcc -arch=sm_100a -ptx oracle/pointnet2_contraction_probe.cu -o - | grep -E "fma|mul.f32|setp.le.f64"
// nvcc 12.9 (same for sm_70): mag = fma(z,z, fma(x,x, y*y)); d = fma(dz,dz, fma(dx,dx, dy*dy)); the 1e-3 test in double.
// oracle_fps_pointnet2 (tokenizer_oracle.c) and the kernel's pointnet2 mode (csrc/fps.cu) use exactly that.
__global__ void k(const float* dataset, float* temp, int old, int n, float* out, int* outi) {
  int tid = threadIdx.x; const int stride = blockDim.x;
  int besti = 0; float best = -1;
  float x1 = dataset[old * 3 + 0];
  float y1 = dataset[old * 3 + 1];
  float z1 = dataset[old * 3 + 2];
  for (int k = tid; k < n; k += stride) {
    float x2, y2, z2;
    x2 = dataset[k * 3 + 0];
    y2 = dataset[k * 3 + 1];
    z2 = dataset[k * 3 + 2];
    float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
    if (mag <= 1e-3) continue;
    float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
    float d2 = min(d, temp[k]);
    temp[k] = d2;
    besti = d2 > best ? k : besti;
    best = d2 > best ? d2 : best;
  }
  out[tid] = best; outi[tid] = besti;
}
