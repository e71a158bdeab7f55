/* Link-time shims so that the reference's chunk path (chunkset.c, chunkset/{mesher,rle,edit}.c,
 * mem.c, event.c) links without GLFW / FastNoise / the in-game shell.  TEST INFRASTRUCTURE ONLY.
 *   ctx_time            - reference: ctx.c:114 (glfwGetTime)
 *   shell_bind_command  - reference: shell.h (console registry, used by mem_init mem.c:261)
 *   noise_*             - reference: cpp/noise.cpp:6-27 (FastNoise, un-vendored).  Never called by
 *                         the harness (worldgen input comes from our own generator). */
#include <time.h>
#include <stdlib.h>

double ctx_time(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void shell_bind_command(char *name, void (*callback)(int, char **))
{
	(void)name; (void)callback;
}

void  noise_init(void) {}
float noise_randf(void) { return 0.0f; }
float noise_simplex(float x, float y, float z) { (void)x; (void)y; (void)z; return 0.0f; }
