/* fake_gl.c -- TEST INFRASTRUCTURE ONLY ("Oracle A" shim).
 *
 * Lets the reference's own host code (/root/reference/src/DoonEngine/voxel.c,
 * compiled IN PLACE by oracle/Makefile, never copied) run without OpenGL:
 *   - the 20 glad_gl* function pointers voxel.c calls are pointed at
 *     host-memory emulations of buffer objects (SSBOs live in malloc memory);
 *   - the reference's shader utility layer (utility/shader.h, out of scope:
 *     there is no run-time shader compilation in a CUDA build) is replaced by
 *     stubs that record uniforms by name;
 *   - glDispatchCompute() executes Oracle B (shader_cpu.c) on the bound
 *     buffers with the recorded uniforms, so DN_draw / DN_update_lighting of
 *     the reference drive the CPU restatement exactly as they would drive the
 *     GLSL programs.
 *
 * Known reference quirk kept visible here: DN_create_volume sizes the map and
 * chunk SSBOs with x*y*x (voxel.c:173,184); maps used with this shim must have
 * mapSize.z <= mapSize.x.
 */
#include <GLAD/glad.h>
#include "DoonEngine/utility/shader.h"
#include "DoonEngine/voxel.h"
#include "oracle.h"

#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------ */
/* buffer objects                                                       */

#define FGL_MAX_BUFFERS 256
#define FGL_MAX_TEXTURES 64
#define FGL_MAX_UNIFORMS 64
#define FGL_PROGRAM_LIGHTING 1
#define FGL_PROGRAM_DRAW 2

typedef struct { int used; size_t size; void* data; } FglBuffer;
typedef struct { int used; int w, h; float* pixels; OrbHit* hits; } FglTexture;
typedef struct { char name[48]; uint32_t words[16]; int numWords; } FglUniform;

static FglBuffer  g_buffers[FGL_MAX_BUFFERS];
static FglTexture g_textures[FGL_MAX_TEXTURES];
static FglUniform g_uniforms[3][FGL_MAX_UNIFORMS];
static int        g_numUniforms[3];

static GLuint g_boundSSBO, g_boundCopyRead, g_boundCopyWrite;
static GLuint g_bindingBase[8];
static GLuint g_boundTexture2D, g_imageTexture;
static GLuint g_currentProgram;
static unsigned g_lastDispatch[3][3];
static int g_execute = 1;
static OrbCounters g_counters[3];
static size_t g_uploadBytes, g_uploadCalls;

static GLuint* bound_slot(GLenum target)
{
	switch(target)
	{
	case GL_COPY_READ_BUFFER:  return &g_boundCopyRead;
	case GL_COPY_WRITE_BUFFER: return &g_boundCopyWrite;
	default:                   return &g_boundSSBO;
	}
}

static FglBuffer* bound_buffer(GLenum target)
{
	GLuint id = *bound_slot(target);
	return (id > 0 && id < FGL_MAX_BUFFERS && g_buffers[id].used) ? &g_buffers[id] : NULL;
}

static void APIENTRY fgl_GenBuffers(GLsizei n, GLuint* out)
{
	for(GLsizei k = 0; k < n; k++)
	{
		out[k] = 0;
		for(GLuint i = 1; i < FGL_MAX_BUFFERS; i++)
			if(!g_buffers[i].used)
			{
				g_buffers[i].used = 1;
				g_buffers[i].size = 0;
				g_buffers[i].data = NULL;
				out[k] = i;
				break;
			}
	}
}

static void APIENTRY fgl_DeleteBuffers(GLsizei n, const GLuint* ids)
{
	for(GLsizei k = 0; k < n; k++)
	{
		GLuint id = ids[k];
		if(id > 0 && id < FGL_MAX_BUFFERS && g_buffers[id].used)
		{
			free(g_buffers[id].data);
			memset(&g_buffers[id], 0, sizeof(FglBuffer));
		}
	}
}

static void APIENTRY fgl_BindBuffer(GLenum target, GLuint id) { *bound_slot(target) = id; }

static void APIENTRY fgl_BindBufferBase(GLenum target, GLuint index, GLuint id)
{
	(void)target;
	if(index < 8)
		g_bindingBase[index] = id;
	g_boundSSBO = id;
}

static void APIENTRY fgl_BufferData(GLenum target, GLsizeiptr size, const void* data, GLenum usage)
{
	(void)usage;
	FglBuffer* b = bound_buffer(target);
	if(!b)
		return;
	free(b->data);
	b->data = calloc((size_t)size + 64, 1);
	b->size = (size_t)size;
	if(data)
		memcpy(b->data, data, (size_t)size);
}

static void APIENTRY fgl_BufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void* data)
{
	FglBuffer* b = bound_buffer(target);
	if(!b || (size_t)offset + (size_t)size > b->size)
		return;
	memcpy((char*)b->data + offset, data, (size_t)size);
	g_uploadBytes += (size_t)size;
	g_uploadCalls++;
}

static void APIENTRY fgl_ClearBufferData(GLenum target, GLenum internalformat, GLenum format, GLenum type, const void* data)
{
	(void)internalformat; (void)format; (void)type; (void)data;
	FglBuffer* b = bound_buffer(target);
	if(b)
		memset(b->data, 0, b->size);
}

static void APIENTRY fgl_CopyBufferSubData(GLenum readTarget, GLenum writeTarget, GLintptr readOffset, GLintptr writeOffset, GLsizeiptr size)
{
	FglBuffer* r = bound_buffer(readTarget);
	FglBuffer* w = bound_buffer(writeTarget);
	if(!r || !w || (size_t)readOffset + (size_t)size > r->size + 64 || (size_t)writeOffset + (size_t)size > w->size + 64)
		return;
	memmove((char*)w->data + writeOffset, (char*)r->data + readOffset, (size_t)size);
}

static void* APIENTRY fgl_MapBuffer(GLenum target, GLenum access)
{
	(void)access;
	FglBuffer* b = bound_buffer(target);
	return b ? b->data : NULL;
}

static GLboolean APIENTRY fgl_UnmapBuffer(GLenum target) { (void)target; return GL_TRUE; }
static GLenum APIENTRY fgl_GetError(void) { return GL_NO_ERROR; }
static void APIENTRY fgl_MemoryBarrier(GLbitfield barriers) { (void)barriers; }
static void APIENTRY fgl_ActiveTexture(GLenum texture) { (void)texture; }

static void APIENTRY fgl_BindTexture(GLenum target, GLuint texture)
{
	if(target == GL_TEXTURE_2D)
		g_boundTexture2D = texture;
}

static void APIENTRY fgl_BindImageTexture(GLuint unit, GLuint texture, GLint level, GLboolean layered, GLint layer, GLenum access, GLenum format)
{
	(void)unit; (void)level; (void)layered; (void)layer; (void)access; (void)format;
	g_imageTexture = texture;
}

static void APIENTRY fgl_GetTexLevelParameteriv(GLenum target, GLint level, GLenum pname, GLint* params)
{
	(void)target; (void)level;
	*params = 0;
	if(g_boundTexture2D > 0 && g_boundTexture2D < FGL_MAX_TEXTURES && g_textures[g_boundTexture2D].used)
		*params = (pname == GL_TEXTURE_WIDTH) ? g_textures[g_boundTexture2D].w : g_textures[g_boundTexture2D].h;
}

/* ------------------------------------------------------------------ */
/* uniforms                                                             */

static FglUniform* uniform_slot(GLuint program, const char* name)
{
	if(program < 1 || program > 2)
		return NULL;
	for(int i = 0; i < g_numUniforms[program]; i++)
		if(strcmp(g_uniforms[program][i].name, name) == 0)
			return &g_uniforms[program][i];
	if(g_numUniforms[program] >= FGL_MAX_UNIFORMS)
		return NULL;
	FglUniform* u = &g_uniforms[program][g_numUniforms[program]++];
	memset(u, 0, sizeof(*u));
	strncpy(u->name, name, sizeof(u->name) - 1);
	return u;
}

static void uniform_store(GLuint program, const char* name, const void* src, int numWords)
{
	FglUniform* u = uniform_slot(program, name);
	if(!u)
		return;
	memcpy(u->words, src, (size_t)numWords * 4);
	u->numWords = numWords;
}

static GLint APIENTRY fgl_GetUniformLocation(GLuint program, const GLchar* name)
{
	FglUniform* u = uniform_slot(program, name);
	return u ? (GLint)(u - g_uniforms[program]) : -1;
}

static void APIENTRY fgl_Uniform3uiv(GLint location, GLsizei count, const GLuint* value)
{
	(void)count;
	if(g_currentProgram < 1 || g_currentProgram > 2 || location < 0 || location >= g_numUniforms[g_currentProgram])
		return;
	memcpy(g_uniforms[g_currentProgram][location].words, value, 12);
	g_uniforms[g_currentProgram][location].numWords = 3;
}

static void APIENTRY fgl_UseProgram(GLuint program) { g_currentProgram = program; }

/* replacements for the reference's utility/shader.c (signatures: utility/shader.h) */
int  DN_compute_program_load(const char* path, const char* includePath)
{
	(void)includePath;
	if(strstr(path, "Lighting")) return FGL_PROGRAM_LIGHTING;
	if(strstr(path, "Draw"))     return FGL_PROGRAM_DRAW;
	return -1;
}
void DN_program_free(GLprogram id) { (void)id; }
void DN_program_activate(GLprogram id) { g_currentProgram = id; }
void DN_program_uniform_int   (GLprogram id, const char* name, GLint    val) { uniform_store(id, name, &val, 1); }
void DN_program_uniform_uint  (GLprogram id, const char* name, GLuint   val) { uniform_store(id, name, &val, 1); }
void DN_program_uniform_float (GLprogram id, const char* name, GLfloat  val) { uniform_store(id, name, &val, 1); }
void DN_program_uniform_double(GLprogram id, const char* name, GLdouble val) { float f = (float)val; uniform_store(id, name, &f, 1); }
void DN_program_uniform_vec2(GLprogram id, const char* name, DNvec2* val) { uniform_store(id, name, val, 2); }
void DN_program_uniform_vec3(GLprogram id, const char* name, DNvec3* val) { uniform_store(id, name, val, 3); }
void DN_program_uniform_vec4(GLprogram id, const char* name, DNvec4* val) { uniform_store(id, name, val, 4); }
void DN_program_uniform_mat3(GLprogram id, const char* name, DNmat3* val) { uniform_store(id, name, val, 9); }
void DN_program_uniform_mat4(GLprogram id, const char* name, DNmat4* val) { uniform_store(id, name, val, 16); }

static void uniform_read(GLuint program, const char* name, void* dst, int numWords)
{
	FglUniform* u = uniform_slot(program, name);
	memset(dst, 0, (size_t)numWords * 4);
	if(u)
		memcpy(dst, u->words, (size_t)(numWords < u->numWords ? numWords : u->numWords) * 4);
}

/* assemble the uniform block of `program` the way the GLSL side would see it */
void fgl_collect_uniforms(int program, OrbUniforms* u)
{
	memset(u, 0, sizeof(*u));
	uniform_read(program, "mapSize", u->mapSize, 3);
	uniform_read(program, "useCubemap", &u->useCubemap, 1);
	uniform_read(program, "skyGradientBot", u->skyGradientBot, 3);
	uniform_read(program, "skyGradientTop", u->skyGradientTop, 3);
	uniform_read(program, "sunStrength", u->sunStrength, 3);
	uniform_read(program, "ambientStrength", u->ambientStrength, 3);
	uniform_read(program, "viewMode", &u->viewMode, 1);
	uniform_read(program, "composeRasterized", &u->composeRasterized, 1);
	uniform_read(program, "invViewMat", u->invViewMat, 16);
	uniform_read(program, "invCenteredViewMat", u->invCenteredViewMat, 16);
	uniform_read(program, "invProjectionMat", u->invProjectionMat, 16);
	uniform_read(program, "time", &u->time, 1);
	uniform_read(program, "numDiffuseSamples", &u->numDiffuseSamples, 1);
	uniform_read(program, "maxDiffuseSamples", &u->maxDiffuseSamples, 1);
	uniform_read(program, "diffuseBounceLimit", &u->diffuseBounceLimit, 1);
	uniform_read(program, "specularBounceLimit", &u->specularBounceLimit, 1);
	uniform_read(program, "sunDir", u->sunDir, 3);
	uniform_read(program, "shadowSoftness", &u->shadowSoftness, 1);
	uniform_read(program, "camPos", u->camPos, 3);
}

/* ------------------------------------------------------------------ */
/* dispatch = run Oracle B, or -- after fgl_set_device -- whatever the harness plugged in: oracle/_ref/libglsl_ref.so's glsl_draw /
 * glsl_light, i.e. the reference's OWN shaders compiled as C++; together with this file's host that is the complete reference
 * pipeline, both halves unrestated (tests/golden/make_golden.py generates the fixtures that way) */
typedef void (*FglDrawFn)(const OrbBuffers*, const OrbUniforms*, int, int, float*, OrbHit*);
typedef void (*FglLightFn)(const OrbBuffers*, const OrbUniforms*, const uint32_t*, size_t, size_t);
static FglDrawFn  g_deviceDraw = NULL;
static FglLightFn g_deviceLight = NULL;

void fgl_set_device(void* drawFn, void* lightFn)
{
	g_deviceDraw = (FglDrawFn)drawFn;
	g_deviceLight = (FglLightFn)lightFn;
}

static void APIENTRY fgl_DispatchCompute(GLuint x, GLuint y, GLuint z)
{
	GLuint p = g_currentProgram;
	if(p < 1 || p > 2)
		return;
	g_lastDispatch[p][0] = x; g_lastDispatch[p][1] = y; g_lastDispatch[p][2] = z;
	if(!g_execute)
		return;

	OrbUniforms u;
	fgl_collect_uniforms((int)p, &u);

	OrbBuffers b;
	b.map       = (OrbHandle*)g_buffers[g_bindingBase[0]].data;
	b.chunks    = (OrbChunk*)g_buffers[g_bindingBase[1]].data;
	b.materials = (const OrbMaterial*)g_buffers[g_bindingBase[2]].data;
	b.voxels    = (OrbVoxel*)g_buffers[g_bindingBase[4]].data;

	if(p == FGL_PROGRAM_LIGHTING)
	{
		const uint32_t* requests = (const uint32_t*)g_buffers[g_bindingBase[3]].data;
		if(g_deviceLight)
			g_deviceLight(&b, &u, requests, x, g_buffers[g_bindingBase[4]].size / sizeof(OrbVoxel));
		else
			orb_light(&b, &u, requests, x, g_buffers[g_bindingBase[4]].size / sizeof(OrbVoxel), &g_counters[p]);
	}
	else
	{
		FglTexture* t = &g_textures[g_imageTexture];
		if(g_imageTexture > 0 && g_imageTexture < FGL_MAX_TEXTURES && t->used)
		{
			if(g_deviceDraw)
				g_deviceDraw(&b, &u, t->w, t->h, t->pixels, t->hits);
			else
				orb_draw(&b, &u, t->w, t->h, t->pixels, t->hits, &g_counters[p]);
		}
	}
}

/* ------------------------------------------------------------------ */
/* harness-facing entry points                                          */

void fgl_init(void)
{
	glad_glActiveTexture = fgl_ActiveTexture;
	glad_glBindBuffer = fgl_BindBuffer;
	glad_glBindBufferBase = fgl_BindBufferBase;
	glad_glBindImageTexture = fgl_BindImageTexture;
	glad_glBindTexture = fgl_BindTexture;
	glad_glBufferData = fgl_BufferData;
	glad_glBufferSubData = fgl_BufferSubData;
	glad_glClearBufferData = fgl_ClearBufferData;
	glad_glCopyBufferSubData = fgl_CopyBufferSubData;
	glad_glDeleteBuffers = fgl_DeleteBuffers;
	glad_glDispatchCompute = fgl_DispatchCompute;
	glad_glGenBuffers = fgl_GenBuffers;
	glad_glGetError = fgl_GetError;
	glad_glGetTexLevelParameteriv = fgl_GetTexLevelParameteriv;
	glad_glGetUniformLocation = fgl_GetUniformLocation;
	glad_glMapBuffer = fgl_MapBuffer;
	glad_glMemoryBarrier = fgl_MemoryBarrier;
	glad_glUniform3uiv = fgl_Uniform3uiv;
	glad_glUnmapBuffer = fgl_UnmapBuffer;
	glad_glUseProgram = fgl_UseProgram;
}

unsigned fgl_create_texture(int w, int h)
{
	for(unsigned i = 1; i < FGL_MAX_TEXTURES; i++)
		if(!g_textures[i].used)
		{
			g_textures[i].used = 1;
			g_textures[i].w = w;
			g_textures[i].h = h;
			g_textures[i].pixels = (float*)calloc((size_t)w * h * 4, sizeof(float));
			g_textures[i].hits = (OrbHit*)calloc((size_t)w * h, sizeof(OrbHit));
			return i;
		}
	return 0;
}

void fgl_delete_texture(unsigned id)
{
	if(id > 0 && id < FGL_MAX_TEXTURES && g_textures[id].used)
	{
		free(g_textures[id].pixels);
		free(g_textures[id].hits);
		memset(&g_textures[id], 0, sizeof(FglTexture));
	}
}

float*  fgl_texture_pixels(unsigned id) { return (id < FGL_MAX_TEXTURES && g_textures[id].used) ? g_textures[id].pixels : NULL; }
OrbHit* fgl_texture_hits(unsigned id)   { return (id < FGL_MAX_TEXTURES && g_textures[id].used) ? g_textures[id].hits : NULL; }
void*   fgl_buffer_ptr(unsigned id)     { return (id < FGL_MAX_BUFFERS && g_buffers[id].used) ? g_buffers[id].data : NULL; }
size_t  fgl_buffer_size(unsigned id)    { return (id < FGL_MAX_BUFFERS && g_buffers[id].used) ? g_buffers[id].size : 0; }
unsigned fgl_binding(unsigned index)    { return index < 8 ? g_bindingBase[index] : 0; }
void    fgl_set_execute(int on)         { g_execute = on; }
void    fgl_last_dispatch(int program, unsigned out[3]) { memcpy(out, g_lastDispatch[program], sizeof(unsigned) * 3); }
void    fgl_get_counters(int program, OrbCounters* out) { *out = g_counters[program]; }
void    fgl_reset_counters(void)        { memset(g_counters, 0, sizeof(g_counters)); g_uploadBytes = 0; g_uploadCalls = 0; }
size_t  fgl_upload_bytes(void)          { return g_uploadBytes; }

/* Emulates what rays do to not-yet-resident tiles (voxelShared.comp:462-466):
 * every tile in state 1 becomes state 3 so the next DN_sync_gpu(DN_WRITE)
 * uploads it ("resident mode" pre-warm). */
size_t fgl_request_all_unloaded(DNvolume* vol)
{
	OrbHandle* map = (OrbHandle*)fgl_buffer_ptr(vol->glMapBufferID);
	size_t n = (size_t)vol->mapSize.x * vol->mapSize.y * vol->mapSize.z, count = 0;
	for(size_t i = 0; i < n; i++)
		if((map[i].flags & 3) == 1)
		{
			map[i].flags = 3;
			count++;
		}
	return count;
}

/* marks every loaded tile visible, as a full-coverage draw would (voxelDraw.comp:121) */
size_t fgl_mark_all_visible(DNvolume* vol)
{
	OrbHandle* map = (OrbHandle*)fgl_buffer_ptr(vol->glMapBufferID);
	size_t n = (size_t)vol->mapSize.x * vol->mapSize.y * vol->mapSize.z, count = 0;
	for(size_t i = 0; i < n; i++)
		if((map[i].flags & 3) == 2)
		{
			map[i].flags |= 4;
			count++;
		}
	return count;
}

size_t fgl_sizeof_volume(void) { return sizeof(DNvolume); }
size_t fgl_sizeof_chunk(void)  { return sizeof(DNchunk); }

/* default message sink; tests may replace g_DN_message_callback */
static char g_lastMessage[512];
static int g_numMessages;
static void fgl_message(DNmessageType type, DNmessageSeverity severity, const char* message)
{
	(void)type; (void)severity;
	strncpy(g_lastMessage, message, sizeof(g_lastMessage) - 1);
	g_numMessages++;
}
void fgl_install_message_sink(void) { g_DN_message_callback = fgl_message; }
const char* fgl_last_message(void)  { return g_lastMessage; }
int fgl_num_messages(void)          { return g_numMessages; }

/* by-pointer wrappers: DNmat4 carries an __m128 member (16-byte aligned, passed
 * by value to DN_draw, voxel.h:207) which ctypes on Python 3.12 cannot express */
void fgl_draw(DNvolume* vol, unsigned tex, const float* view, const float* projection)
{
	DNmat4 v, p;
	memcpy(&v, view, sizeof(DNmat4));
	memcpy(&p, projection, sizeof(DNmat4));
	DN_draw(vol, tex, v, p, -1, -1);
}

void fgl_set_view_projection(DNvolume* vol, float aspect, float nearPlane, float farPlane, float* view, float* projection)
{
	DNmat4 v, p;
	DN_set_view_projection_matrices(vol, aspect, nearPlane, farPlane, &v, &p);
	memcpy(view, &v, sizeof(DNmat4));
	memcpy(projection, &p, sizeof(DNmat4));
}

/* bulk form of 512 DN_set_compressed_voxel calls (the reference has no bulk entry point); voxels[x][y][z][2] */
void fgl_set_chunk(DNvolume* vol, int mx, int my, int mz, const uint32_t* voxels)
{
	DNivec3 mapPos = {mx, my, mz};
	for(int x = 0; x < 8; x++)
		for(int y = 0; y < 8; y++)
			for(int z = 0; z < 8; z++)
			{
				const uint32_t* w = voxels + ((x * 8 + y) * 8 + z) * 2;
				DNcompressedVoxel v = {w[0], w[1]};
				DN_set_compressed_voxel(vol, mapPos, (DNivec3){x, y, z}, v);
			}
}
