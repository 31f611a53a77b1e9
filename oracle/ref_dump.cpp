/* ref_dump.cpp — build-time helper of oracle/build_ref.py (TEST INFRASTRUCTURE ONLY).
 *
 * Compiled with -I/root/reference/gsplat_plugin/shaders, i.e. against the reference's own headers WHERE THEY LIE
 * (GSplatShaderSource.h includes GSplatShaderCoreLib.h; both need only <string>).  It writes the reference's shader
 * strings, byte for byte, into the scratch directory given on the command line, where build_ref.py turns them into
 * the include files of ref_harness.cpp.  Nothing it writes is kept: only oracle/_ref/libgsplat_ref.so survives the build.
 */
#include <cstdio>
#include <initializer_list>
#include <string>

#include "GSplatShaderSource.h"   /* /root/reference/gsplat_plugin/shaders (via -I) */

static int dump(const std::string& dir, const char* name, const char* text)
{
    const std::string path = dir + "/" + name;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); return 1; }
    fputs(text, f);
    fclose(f);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: ref_dump <scratch dir>\n"); return 2; }
    const std::string dir = argv[1];
    int rc = 0;
    rc |= dump(dir, "core_lib.glsl", GSplatCoreLib);                       /* GSplatShaderCoreLib.h:8-95   */
    rc |= dump(dir, "sh_lib.glsl", GSplatSphericalHarmonicsLib);           /* GSplatShaderCoreLib.h:101-181 */
    rc |= dump(dir, "main_vs.glsl", _GSplatMainVertexShader);              /* GSplatShaderSource.h:117-288 */
    rc |= dump(dir, "main_fs.glsl", _GSplatMainFragmentShader);            /* GSplatShaderSource.h:291-314 */
    rc |= dump(dir, "wire_vs.glsl", _GSplatWireVertexShader);              /* GSplatShaderSource.h:22-90   */
    rc |= dump(dir, "wire_fs.glsl", _GSplatWireFragmentShader);            /* GSplatShaderSource.h:93-110  */
    return rc;
}
