/*
 * slharness.c -- MEASUREMENT HARNESS, not part of the product library.
 *
 * The reference's programs are C++: a frame is a loop of `uniforms.assign(...); context.draw(mesh, shader, fbo);`
 * (softlight/tests/sl_large_scene_test.cpp:700-730, sl_animation_test.cpp).  The test / bench harness of this repo is
 * Python, where each such call costs about a microsecond of interpreter time -- 0.8 ms for the 768 calls of the indoor
 * scene's frame.  This file replays a recorded list of those calls from C, through whichever library's C ABI entry points
 * the caller passes in (the CUDA product, the compiled reference or the CPU restatement), so that a measured frame
 * contains the call sequence a C/C++ host would issue and nothing else.
 */
#include <stddef.h>
#include <stdint.h>

typedef struct slh_cmd
{
    int32_t op; /* 0 ubo_assign, 1 ubo_bind_texture, 2 ubo_bind_buffer, 3 draw, 4 draw_instanced */
    int32_t pad_;
    uint64_t a, b, c, d;
    const void* p;
} slh_cmd;

typedef struct slh_api
{
    int (*ubo_assign)(void* ctx, size_t ubo, const void* data, size_t offset, size_t numBytes);
    int (*ubo_bind_texture)(void* ctx, size_t ubo, size_t offset, size_t texture);
    int (*ubo_bind_buffer)(void* ctx, size_t ubo, size_t offset, size_t vbo, size_t byteOffset);
    int (*draw)(void* ctx, const void* mesh, size_t shader, size_t fbo);
    int (*draw_instanced)(void* ctx, const void* mesh, size_t numInstances, size_t shader, size_t fbo);
} slh_api;

/* returns 0, or the failing call's index + 1 in the upper bits and its return code in *rcOut */
size_t slh_replay(const slh_api* api, void* ctx, const slh_cmd* cmds, size_t n, int* rcOut)
{
    for (size_t i = 0; i < n; ++i)
    {
        const slh_cmd* k = &cmds[i];
        int rc = 0;
        switch (k->op)
        {
            case 0: rc = api->ubo_assign(ctx, (size_t)k->a, k->p, (size_t)k->b, (size_t)k->c); break;
            case 1: rc = api->ubo_bind_texture(ctx, (size_t)k->a, (size_t)k->b, (size_t)k->c); break;
            case 2: rc = api->ubo_bind_buffer(ctx, (size_t)k->a, (size_t)k->b, (size_t)k->c, (size_t)k->d); break;
            case 3: rc = api->draw(ctx, k->p, (size_t)k->a, (size_t)k->b); break;
            case 4: rc = api->draw_instanced(ctx, k->p, (size_t)k->a, (size_t)k->b, (size_t)k->c); break;
            default: rc = -1000; break;
        }
        if (rc < 0)
        {
            if (rcOut) *rcOut = rc;
            return i + 1;
        }
    }
    if (rcOut) *rcOut = 0;
    return 0;
}
