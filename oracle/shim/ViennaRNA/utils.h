/* Shim: stem_kernel_lite/data.h includes this inside namespace Vienna { extern "C" {...} }
 * and only calls init_rand(); the harness defines it as a no-op. */
void init_rand(void);
