/* oracle/gfx_shim/gl_capture.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A headless stand-in for the OpenGL entry points the reference's gfx/vsplat.c calls, so that its gfx_update_svl
 * (gfx/vsplat.c:197-338) -- the consumer right after the chunk-rebuild path, SURVEY 8(f) row f2 -- can run UNMODIFIED and
 * what it uploads can be read back: buffer objects are plain host allocations (glBufferData / glBufferSubData are
 * recorded), everything else is a no-op.  glad.h declares every gl* name as a function pointer `glad_gl*`; the ones
 * vsplat.c references are defined here.  Nothing of the reference is copied. */
#include <glad/glad.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define CAP_MAX_BUFFERS (1u << 20)
static struct { void *data; size_t size; } cap_buf[CAP_MAX_BUFFERS];
static GLuint cap_next_buffer = 1, cap_next_vao = 1, cap_bound_array;

static void APIENTRY cap_GenBuffers(GLsizei n, GLuint *ids) { for (GLsizei i = 0; i < n; i++) ids[i] = cap_next_buffer < CAP_MAX_BUFFERS ? cap_next_buffer++ : 0; }
static void APIENTRY cap_DeleteBuffers(GLsizei n, const GLuint *ids)
{
	for (GLsizei i = 0; i < n; i++) if (ids[i] && ids[i] < CAP_MAX_BUFFERS) { free(cap_buf[ids[i]].data); cap_buf[ids[i]].data = NULL; cap_buf[ids[i]].size = 0; }
}
static void APIENTRY cap_BindBuffer(GLenum target, GLuint id) { if (target == GL_ARRAY_BUFFER) cap_bound_array = id; }
static void APIENTRY cap_BufferData(GLenum target, GLsizeiptr size, const void *data, GLenum usage)
{
	(void)usage;
	if (target != GL_ARRAY_BUFFER || !cap_bound_array || cap_bound_array >= CAP_MAX_BUFFERS) return;
	free(cap_buf[cap_bound_array].data);
	cap_buf[cap_bound_array].data = calloc((size_t)size + 1, 1);
	cap_buf[cap_bound_array].size = (size_t)size;
	if (data) memcpy(cap_buf[cap_bound_array].data, data, (size_t)size);
}
static void APIENTRY cap_BufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void *data)
{
	if (target != GL_ARRAY_BUFFER || !cap_bound_array || cap_bound_array >= CAP_MAX_BUFFERS) return;
	if ((size_t)offset + (size_t)size > cap_buf[cap_bound_array].size) abort();          /* a GL_INVALID_VALUE in a real context */
	memcpy((uint8_t *)cap_buf[cap_bound_array].data + offset, data, (size_t)size);
}
static void APIENTRY cap_GenVertexArrays(GLsizei n, GLuint *ids) { for (GLsizei i = 0; i < n; i++) ids[i] = cap_next_vao++; }
static void APIENTRY cap_DeleteVertexArrays(GLsizei n, const GLuint *ids) { (void)n; (void)ids; }
static void APIENTRY cap_u1(GLuint a) { (void)a; }
static void APIENTRY cap_e1(GLenum a) { (void)a; }
static void APIENTRY cap_VertexAttribPointer(GLuint i, GLint s, GLenum t, GLboolean nrm, GLsizei st, const void *p) { (void)i; (void)s; (void)t; (void)nrm; (void)st; (void)p; }
static void APIENTRY cap_VertexAttribIPointer(GLuint i, GLint s, GLenum t, GLsizei st, const void *p) { (void)i; (void)s; (void)t; (void)st; (void)p; }

PFNGLGENBUFFERSPROC glad_glGenBuffers = cap_GenBuffers;
PFNGLDELETEBUFFERSPROC glad_glDeleteBuffers = cap_DeleteBuffers;
PFNGLBINDBUFFERPROC glad_glBindBuffer = cap_BindBuffer;
PFNGLBUFFERDATAPROC glad_glBufferData = cap_BufferData;
PFNGLBUFFERSUBDATAPROC glad_glBufferSubData = cap_BufferSubData;
PFNGLGENVERTEXARRAYSPROC glad_glGenVertexArrays = cap_GenVertexArrays;
PFNGLDELETEVERTEXARRAYSPROC glad_glDeleteVertexArrays = cap_DeleteVertexArrays;
PFNGLBINDVERTEXARRAYPROC glad_glBindVertexArray = cap_u1;
PFNGLUSEPROGRAMPROC glad_glUseProgram = cap_u1;
PFNGLENABLEVERTEXATTRIBARRAYPROC glad_glEnableVertexAttribArray = cap_u1;
PFNGLVERTEXATTRIBPOINTERPROC glad_glVertexAttribPointer = cap_VertexAttribPointer;
PFNGLVERTEXATTRIBIPOINTERPROC glad_glVertexAttribIPointer = cap_VertexAttribIPointer;
/* referenced by the init / draw code of vsplat.c, which the harness never runs */
PFNGLACTIVETEXTUREPROC glad_glActiveTexture = cap_e1;
PFNGLBINDTEXTUREPROC glad_glBindTexture;
PFNGLBLENDFUNCPROC glad_glBlendFunc;
PFNGLDISABLEPROC glad_glDisable = cap_e1;
PFNGLENABLEPROC glad_glEnable = cap_e1;
PFNGLDRAWARRAYSPROC glad_glDrawArrays;
PFNGLGENTEXTURESPROC glad_glGenTextures;
PFNGLGENERATEMIPMAPPROC glad_glGenerateMipmap;
PFNGLGETATTRIBLOCATIONPROC glad_glGetAttribLocation;
PFNGLGETUNIFORMLOCATIONPROC glad_glGetUniformLocation;
PFNGLTEXIMAGE2DPROC glad_glTexImage2D;
PFNGLTEXPARAMETERIPROC glad_glTexParameteri;
PFNGLUNIFORM1FPROC glad_glUniform1f;
PFNGLUNIFORM2FPROC glad_glUniform2f;
PFNGLUNIFORMMATRIX4FVPROC glad_glUniformMatrix4fv;

/* engine symbols of the same unused code */
void ctx_get_window_size(int *w, int *h) { *w = 0; *h = 0; }
unsigned int gfx_create_shader(const char *v, const char *f) { (void)v; (void)f; return 1; }
void *res_file(int id) { (void)id; return NULL; }
unsigned long res_size(int id) { (void)id; return 0; }
char *res_strcpy(int id) { (void)id; return NULL; }
unsigned char *stbi_load_from_memory(const unsigned char *b, int l, int *x, int *y, int *c, int d) { (void)b; (void)l; (void)x; (void)y; (void)c; (void)d; return NULL; }

/* read-back for the harness */
const void *vr_gl_buffer(unsigned int id, size_t *size)
{
	if (!id || id >= CAP_MAX_BUFFERS) { *size = 0; return NULL; }
	*size = cap_buf[id].size;
	return cap_buf[id].data;
}
