/* include/hpfw_b200_pyhpfw.h — the reference's OWN C ABI, re-exported by libhpfw_b200.so.
 *
 * Replaces /root/reference/modules/python/parallel_collector_wrapper.hpp:12-38 (implementation .cpp:5-62): the same eight
 * symbols with the same struct layout and ownership rules, so modules/python/pyhpfw/pyhpfw.py (ctypes) works against this
 * library by changing only the path it dlopens. Behind the handle is hpfw::ParallelCollector<HashprintHandle<uint64_t,
 * CQT<>, 20, 80>, cache::DriveCache> from include/hpfw/, i.e. the GPU path.
 *
 * Ownership: results are allocated with new[] by the callee and must be released with the matching *_free call.
 * Errors: the reference lets C++ exceptions unwind through extern "C" (undefined behaviour). Here every entry point
 * catches: on failure it returns NULL (and stores 0 in *got / *size) and hpfw_last_error() holds the message;
 * prepare() skips unreadable files like the reference, so *got may be smaller than n.
 * `cache` argument of save/load: ignored by the reference (.cpp:56-62); here a non-empty string re-targets the cache
 * directory, NULL or "" keeps "cache/".
 */
#ifndef HPFW_B200_PYHPFW_H
#define HPFW_B200_PYHPFW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct FilenameHashprintPair {   /* parallel_collector_wrapper.hpp:12-16 */
    char *filename;
    uint64_t *hashprint;
    int hp_size;
} FilenameHashprintPair;

typedef struct LiveIdCollector LiveIdCollector;

LiveIdCollector *par_collector_new(void);
void par_collector_del(LiveIdCollector *collector);
FilenameHashprintPair *par_collector_prepare(LiveIdCollector *collector, const char **filenames, int n, int *got);
uint64_t *par_collector_calc_hashprint(LiveIdCollector *collector, const char *filename, int *size);
void par_collector_save(LiveIdCollector *collector, const char *cache);
void par_collector_load(LiveIdCollector *collector, const char *cache);
void prepare_result_free(FilenameHashprintPair *res, int got);
void calc_hashprint_result_free(uint64_t *hp);

#ifdef __cplusplus
}
#endif
#endif /* HPFW_B200_PYHPFW_H */
