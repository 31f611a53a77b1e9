#pragma once
/* stub: the OpenGL 3.3 core entry points the shim's two GL helpers call (and the typedefs cuda_gl_interop.h needs) */
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef int GLsizei; typedef float GLfloat;
typedef unsigned char GLboolean; typedef char GLchar; typedef ptrdiff_t GLsizeiptr; typedef ptrdiff_t GLintptr; typedef unsigned int GLbitfield;
#define GL_TEXTURE_2D 0x0DE1
#define GL_TEXTURE0 0x84C0
#define GL_TEXTURE_BINDING_2D 0x8069
#define GL_ACTIVE_TEXTURE 0x84E0
#define GL_CURRENT_PROGRAM 0x8B8D
#define GL_VERTEX_ARRAY_BINDING 0x85B5
#define GL_PIXEL_PACK_BUFFER 0x88EB
#define GL_PIXEL_UNPACK_BUFFER 0x88EC
#define GL_PIXEL_PACK_BUFFER_BINDING 0x88ED
#define GL_PIXEL_UNPACK_BUFFER_BINDING 0x88EF
#define GL_STREAM_COPY 0x88E2
#define GL_DEPTH_COMPONENT 0x1902
#define GL_RED 0x1903
#define GL_FLOAT 0x1406
#define GL_TRIANGLES 0x0004
#define GL_VERTEX_SHADER 0x8B31
#define GL_FRAGMENT_SHADER 0x8B30
extern "C" {
void glGetIntegerv(GLenum, GLint*); void glBindTexture(GLenum, GLuint); void glActiveTexture(GLenum);
void glGenBuffers(GLsizei, GLuint*); void glBindBuffer(GLenum, GLuint); void glBufferData(GLenum, GLsizeiptr, const void*, GLenum);
void glReadPixels(GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, void*);
void glTexSubImage2D(GLenum, GLint, GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, const void*);
GLuint glCreateShader(GLenum); void glShaderSource(GLuint, GLsizei, const GLchar* const*, const GLint*); void glCompileShader(GLuint);
GLuint glCreateProgram(void); void glAttachShader(GLuint, GLuint); void glLinkProgram(GLuint); void glUseProgram(GLuint);
GLint glGetUniformLocation(GLuint, const GLchar*); void glUniform1i(GLint, GLint);
void glGenVertexArrays(GLsizei, GLuint*); void glBindVertexArray(GLuint); void glDrawArrays(GLenum, GLint, GLsizei);
}
