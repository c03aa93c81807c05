/* The drop-in boundary from plain C: links against libskoots_b200.so with nothing but include/skoots_b200.h.
 * Without a GPU it exercises the host side only (version, workspace arithmetic, argument validation);
 * tests/test_host_logic.py builds and runs it.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_check.c -Lskoots_b200 -lskoots_b200 -Wl,-rpath,$PWD/skoots_b200 -o /tmp/c_abi_check
 */
#include <stdio.h>
#include <string.h>

#include "skoots_b200.h"

int main(void) {
    float scale[3] = {60.f, 60.f, 12.f};
    int rc;
    size_t bytes;
    if (skb_version() != SKB_VERSION) {
        printf("header %d != library %d\n", SKB_VERSION, skb_version());
        return 1;
    }
    bytes = skb_ccl_workspace_bytes(2048, 2048, 512, 1 << 20);
    if (bytes < (size_t)2048 * 2048 * 512 / 8) {
        printf("workspace smaller than the bit mask: %zu\n", bytes);
        return 1;
    }
    /* NULL volume: rejected with a code and a message, nothing is enqueued */
    rc = skb_vec_embed3d(NULL, SKB_F16, 1, 8, 8, 8, scale, 1, 1.0, NULL, NULL);
    if (rc == SKB_OK || strlen(skb_last_error()) == 0) {
        printf("a NULL volume was accepted\n");
        return 1;
    }
    printf("skb %d: workspace for 2048x2048x512 = %zu bytes; NULL argument -> %d (%s)\n", skb_version(), bytes, rc,
           skb_last_error());
    return 0;
}
