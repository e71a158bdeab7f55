/* oracle/gfx_shim/stb_image.h -- TEST INFRASTRUCTURE ONLY: declarations for the texture loader gfx_vsplat_init calls
 * (never run by the harness). */
#pragma once
unsigned char *stbi_load_from_memory(const unsigned char *buffer, int len, int *x, int *y, int *channels_in_file, int desired_channels);
void stbi_image_free(void *p);
