// ORACLE-ONLY (test infrastructure): POSIX implementation of the reference's platform
// interface (Raytracer/platform.h:5-35) replacing platform_win32.inl, so that the
// reference's arena, thread pool and timers run headless on Linux.
#include "precomp.h"
#include "common.h"

#include <sys/mman.h>
#include <unistd.h>
#include <time.h>
#include <pthread.h>
#include <semaphore.h>

usize g_platform_page_size;

void platform_init() {
    g_platform_page_size = (usize)sysconf(_SC_PAGESIZE);
}

void* platform_allocate(usize size) {
    void* p = mmap(0, size, PROT_READ|PROT_WRITE, MAP_PRIVATE|MAP_ANONYMOUS, -1, 0);
    return p == MAP_FAILED ? 0 : p;
}

void* platform_reserve(usize size) {
    void* p = mmap(0, size, PROT_NONE, MAP_PRIVATE|MAP_ANONYMOUS|MAP_NORESERVE, -1, 0);
    return p == MAP_FAILED ? 0 : p;
}

b32 platform_commit(void* address, usize size) {
    return mprotect(address, size, PROT_READ|PROT_WRITE) == 0;
}

void platform_free(void* memory) { (void)memory; }

char* platform_read_entire_file(Arena* arena, const char* file_name, usize* out_file_size) {
    char* result = 0;
    FILE* f = fopen(file_name, "rb");
    if (f) {
        fseek(f, 0, SEEK_END);
        long size = ftell(f);
        fseek(f, 0, SEEK_SET);
        result = push_array(arena, (usize)size + 1, char, no_clear());
        size_t got = fread(result, 1, (size_t)size, f);
        result[got] = 0;
        if (out_file_size) *out_file_size = got;
        fclose(f);
    }
    return result;
}

b32 platform_write_entire_file(const char* file_name, usize size, const void* data) {
    FILE* f = fopen(file_name, "wb");
    if (!f) return false;
    size_t put = fwrite(data, 1, size, f);
    fclose(f);
    return put == size;
}

PlatformHighResTime platform_get_timestamp() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    PlatformHighResTime r;
    r.data_ = (u64)ts.tv_sec*1000000000ull + (u64)ts.tv_nsec;
    return r;
}

f64 platform_get_seconds_elapsed(PlatformHighResTime start, PlatformHighResTime end) {
    return (f64)(end.data_ - start.data_)*1e-9;
}

PlatformSemaphore platform_create_semaphore(u32 start_count, u32 max_count) {
    (void)max_count;
    sem_t* s = (sem_t*)malloc(sizeof(sem_t));
    sem_init(s, 0, start_count);
    PlatformSemaphore r; r.opaque = s;
    return r;
}

void platform_destroy_semaphore(PlatformSemaphore semaphore) {
    sem_destroy((sem_t*)semaphore.opaque);
    free(semaphore.opaque);
}

void platform_release_semaphore(PlatformSemaphore semaphore, u32 count, u32* previous_count) {
    sem_t* s = (sem_t*)semaphore.opaque;
    int prev = 0;
    sem_getvalue(s, &prev);
    for (u32 i = 0; i < count; ++i) sem_post(s);
    if (previous_count) *previous_count = (u32)(prev < 0 ? 0 : prev);
}

void platform_wait_on_semaphore(PlatformSemaphore semaphore) {
    sem_t* s = (sem_t*)semaphore.opaque;
    while (sem_wait(s) != 0) {}
}

struct PosixThreadParameters {
    PlatformSemaphore semaphore;
    PlatformThreadProc proc;
    void* userdata;
};

static void* posix_thread_proc(void* p) {
    PosixThreadParameters params = *(PosixThreadParameters*)p;
    params.proc(params.userdata, params.semaphore);
    return 0;
}

b32 platform_create_thread(PlatformThreadProc proc, void* userdata) {
    static PlatformSemaphore thread_creation_semaphore = platform_create_semaphore(0, 1);
    PosixThreadParameters params;
    params.semaphore = thread_creation_semaphore;
    params.proc = proc;
    params.userdata = userdata;
    pthread_t t;
    if (pthread_create(&t, 0, posix_thread_proc, &params) != 0) return false;
    pthread_detach(t);
    platform_wait_on_semaphore(thread_creation_semaphore);
    return true;
}
