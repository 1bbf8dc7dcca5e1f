/* lum_utils.c - the "extra utilities" of Luminary's public API (include/luminary/{array,host_memory,log,queue,ringbuffer,
 * thread_status,name_strings}.h), which applications written against the reference - its own command line front end first of all
 * (src/mandarin_duck/main.c, argument_parser.c, mandarin_duck.c) - call next to the luminary_host_* functions.
 *
 * Own implementations of the documented behaviour (reference: src/luminary/array.c, host_memory.c, log.c, queue.c, ringbuffer.c,
 * thread_status.c, name_strings.c): same names, argument meanings and result codes, nothing else shared. */
#define LUMINARY_INCLUDE_EXTRA_UTILS
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lum_host_internal.h"

/* ------------------------------------------------------------------------------------------------------------ */
/* tracked host memory: every block carries its size so that a leak total can be reported at shutdown            */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct MemHeader {
  uint64_t magic;
  uint64_t size;
} MemHeader;
#define MEM_MAGIC 0x4C554D4232303048ull /* "LUMB200H" */

static pthread_mutex_t g_mem_lock = PTHREAD_MUTEX_INITIALIZER;
static uint64_t g_mem_bytes       = 0;

uint64_t lum_host_memory_in_use(void) {
  pthread_mutex_lock(&g_mem_lock);
  const uint64_t v = g_mem_bytes;
  pthread_mutex_unlock(&g_mem_lock);
  return v;
}

LuminaryResult _host_malloc(void** ptr, size_t size, const char* buf_name, const char* func, uint32_t line) {
  if (!ptr)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "host_malloc(%s) in %s:%u: destination is NULL", buf_name ? buf_name : "?", func ? func : "?", line);
  MemHeader* h = (MemHeader*) malloc(sizeof(MemHeader) + size);
  if (!h) {
    *ptr = NULL;
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "host_malloc(%s, %zu bytes) in %s:%u failed", buf_name ? buf_name : "?", size, func ? func : "?", line);
  }
  h->magic = MEM_MAGIC;
  h->size  = size;
  pthread_mutex_lock(&g_mem_lock);
  g_mem_bytes += size;
  pthread_mutex_unlock(&g_mem_lock);
  *ptr = (void*) (h + 1);
  return LUMINARY_SUCCESS;
}

LuminaryResult _host_realloc(void** ptr, size_t size, const char* buf_name, const char* func, uint32_t line) {
  if (!ptr)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "host_realloc(%s) in %s:%u: destination is NULL", buf_name ? buf_name : "?", func ? func : "?", line);
  if (!*ptr)
    return _host_malloc(ptr, size, buf_name, func, line);
  MemHeader* h = ((MemHeader*) *ptr) - 1;
  if (h->magic != MEM_MAGIC)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "host_realloc(%s) in %s:%u: the pointer was not allocated by host_malloc", buf_name ? buf_name : "?",
                     func ? func : "?", line);
  const uint64_t old = h->size;
  MemHeader* n       = (MemHeader*) realloc(h, sizeof(MemHeader) + size);
  if (!n)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "host_realloc(%s, %zu bytes) in %s:%u failed", buf_name ? buf_name : "?", size, func ? func : "?", line);
  n->size = size;
  pthread_mutex_lock(&g_mem_lock);
  g_mem_bytes += size;
  g_mem_bytes -= old;
  pthread_mutex_unlock(&g_mem_lock);
  *ptr = (void*) (n + 1);
  return LUMINARY_SUCCESS;
}

LuminaryResult _host_free(void** ptr, const char* buf_name, const char* func, uint32_t line) {
  if (!ptr)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "host_free(%s) in %s:%u: pointer address is NULL", buf_name ? buf_name : "?", func ? func : "?", line);
  if (!*ptr)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "host_free(%s) in %s:%u: pointer is NULL", buf_name ? buf_name : "?", func ? func : "?", line);
  MemHeader* h = ((MemHeader*) *ptr) - 1;
  if (h->magic != MEM_MAGIC)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "host_free(%s) in %s:%u: the pointer was not allocated by host_malloc", buf_name ? buf_name : "?",
                     func ? func : "?", line);
  pthread_mutex_lock(&g_mem_lock);
  g_mem_bytes -= h->size;
  pthread_mutex_unlock(&g_mem_lock);
  h->magic = 0;
  free(h);
  *ptr = NULL;
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* arrays: the caller holds a typed pointer to element 0; the bookkeeping sits directly in front of it           */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct ArrayHeader {
  uint64_t magic;
  uint64_t size_of_element;
  uint32_t num_elements;
  uint32_t capacity;
} ArrayHeader;
#define ARRAY_MAGIC 0x4C554D4172726179ull /* "LUMArray" */

static ArrayHeader* array_header(const void* array) { return array ? ((ArrayHeader*) array) - 1 : NULL; }

#define ARRAY_CHECK(h)                                                                      \
  do {                                                                                      \
    if (!(h))                                                                               \
      LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array is NULL");                      \
    if ((h)->magic != ARRAY_MAGIC)                                                          \
      LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Given pointer is not an array.");     \
  } while (0)

LuminaryResult _array_create(void** array, size_t size_of_element, uint32_t num_elements, const char* buf_name, const char* func, uint32_t line) {
  if (!array)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array_create(%s) in %s:%u: destination is NULL", buf_name ? buf_name : "?", func ? func : "?", line);
  if (size_of_element == 0)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "array_create(%s): elements of size 0", buf_name ? buf_name : "?");
  void* block = NULL;
  LUM_TRY(_host_malloc(&block, sizeof(ArrayHeader) + size_of_element * (size_t) num_elements, buf_name, func, line));
  ArrayHeader* h     = (ArrayHeader*) block;
  h->magic           = ARRAY_MAGIC;
  h->size_of_element = size_of_element;
  h->num_elements    = 0;
  h->capacity        = num_elements;
  *array             = (void*) (h + 1);
  return LUMINARY_SUCCESS;
}

LuminaryResult _array_resize(void** array, size_t num_elements, const char* buf_name, const char* func, uint32_t line) {
  if (!array)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array address is NULL");
  ArrayHeader* h = array_header(*array);
  ARRAY_CHECK(h);
  if (num_elements > 0xFFFFFFFFull)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Array exceeded maximum number of elements.");
  void* block = (void*) h;
  LUM_TRY(_host_realloc(&block, sizeof(ArrayHeader) + h->size_of_element * num_elements, buf_name, func, line));
  h           = (ArrayHeader*) block;
  h->capacity = (uint32_t) num_elements;
  if (h->num_elements > h->capacity)
    h->num_elements = h->capacity;
  *array = (void*) (h + 1);
  return LUMINARY_SUCCESS;
}

LuminaryResult _array_destroy(void** array, const char* buf_name, const char* func, uint32_t line) {
  if (!array)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array address is NULL");
  ArrayHeader* h = array_header(*array);
  ARRAY_CHECK(h);
  h->magic    = 0;
  void* block = (void*) h;
  LUM_TRY(_host_free(&block, buf_name, func, line));
  *array = NULL;
  return LUMINARY_SUCCESS;
}

LuminaryResult _array_push(void** array, void* object, const char* buf_name, const char* func, uint32_t line) {
  if (!array || !object)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array_push: NULL argument");
  ArrayHeader* h = array_header(*array);
  ARRAY_CHECK(h);
  if (h->num_elements == 0xFFFFFFFFu)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Array exceeded maximum number of elements.");
  if (h->num_elements == h->capacity) {
    const uint64_t grown = h->capacity ? 2ull * h->capacity : 4ull;
    LUM_TRY(_array_resize(array, (size_t) (grown > 0xFFFFFFFFull ? 0xFFFFFFFFull : grown), buf_name, func, line));
    h = array_header(*array);
  }
  memcpy((uint8_t*) *array + h->size_of_element * h->num_elements, object, h->size_of_element);
  h->num_elements++;
  return LUMINARY_SUCCESS;
}

LuminaryResult _array_append(void** dst, const void* src, const char* buf_name, const char* func, uint32_t line) {
  if (!dst)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array address is NULL");
  ArrayHeader* d       = array_header(*dst);
  const ArrayHeader* s = array_header(src);
  ARRAY_CHECK(s);
  ARRAY_CHECK(d);
  if (d->size_of_element != s->size_of_element)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Array elements are of different size.");
  const uint64_t total = (uint64_t) d->num_elements + s->num_elements;
  if (total > 0xFFFFFFFFull)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Array exceeded maximum number of elements.");
  if (d->capacity < total) {
    LUM_TRY(_array_resize(dst, (size_t) total, buf_name, func, line));
    d = array_header(*dst);
  }
  memcpy((uint8_t*) *dst + d->size_of_element * d->num_elements, src, s->size_of_element * s->num_elements);
  d->num_elements = (uint32_t) total;
  return LUMINARY_SUCCESS;
}

LuminaryResult _array_copy(void** dst, const void* src, const char* buf_name, const char* func, uint32_t line) {
  if (!dst)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array address is NULL");
  ArrayHeader* d = array_header(*dst);
  ARRAY_CHECK(d);
  d->num_elements = 0;
  return _array_append(dst, src, buf_name, func, line);
}

LuminaryResult array_clear(void* array) {
  ArrayHeader* h = array_header(array);
  ARRAY_CHECK(h);
  h->num_elements = 0;
  return LUMINARY_SUCCESS;
}

LuminaryResult array_get_size(const void* array, size_t* size) {
  if (!size)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "size is NULL");
  const ArrayHeader* h = array_header(array);
  ARRAY_CHECK(h);
  *size = h->size_of_element * h->num_elements;
  return LUMINARY_SUCCESS;
}

LuminaryResult array_get_num_elements(const void* array, uint32_t* num_elements) {
  if (!num_elements)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "num_elements is NULL");
  const ArrayHeader* h = array_header(array);
  ARRAY_CHECK(h);
  *num_elements = h->num_elements;
  return LUMINARY_SUCCESS;
}

LuminaryResult _array_set_num_elements(void** array, uint32_t num_elements, const char* buf_name, const char* func, uint32_t line) {
  if (!array)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "array address is NULL");
  ArrayHeader* h = array_header(*array);
  ARRAY_CHECK(h);
  if (num_elements > h->capacity) {
    LUM_TRY(_array_resize(array, num_elements, buf_name, func, line));
    h = array_header(*array);
  }
  if (num_elements > h->num_elements) /* new elements read as zero */
    memset((uint8_t*) *array + h->size_of_element * h->num_elements, 0, h->size_of_element * (num_elements - h->num_elements));
  h->num_elements = num_elements;
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* log: messages go to the console (coloured by severity) and into a memory buffer that luminary_write_log dumps  */
/* ------------------------------------------------------------------------------------------------------------ */
static pthread_mutex_t g_log_lock = PTHREAD_MUTEX_INITIALIZER;
static char* g_log                = NULL;
static size_t g_log_len = 0, g_log_cap = 0;
static bool g_inline_pending = false;

static void log_append(const char* prefix, const char* fmt, va_list ap) {
  char line[4096];
  const int n = vsnprintf(line, sizeof(line), fmt, ap);
  if (n < 0)
    return;
  const size_t len  = (size_t) (n < (int) sizeof(line) ? n : (int) sizeof(line) - 1);
  const size_t plen = strlen(prefix);
  if (g_log_len + plen + len + 2 > g_log_cap) {
    const size_t cap = (g_log_cap ? 2 * g_log_cap : 1 << 16) + plen + len + 2;
    char* grown      = (char*) realloc(g_log, cap);
    if (!grown)
      return;
    g_log     = grown;
    g_log_cap = cap;
  }
  memcpy(g_log + g_log_len, prefix, plen);
  memcpy(g_log + g_log_len + plen, line, len);
  g_log_len += plen + len;
  g_log[g_log_len++] = '\n';
  g_log[g_log_len]   = '\0';
}

static void console(FILE* f, const char* colour, const char* fmt, va_list ap, bool newline) {
  if (g_inline_pending) {
    fputs("\r\033[K", stdout);
    g_inline_pending = false;
  }
  if (colour)
    fputs(colour, f);
  vfprintf(f, fmt, ap);
  if (colour)
    fputs("\033[0m", f);
  if (newline)
    fputc('\n', f);
  fflush(f);
}

void luminary_print_log(const char* format, ...) {
  va_list ap;
  va_start(ap, format);
  pthread_mutex_lock(&g_log_lock);
  log_append("[LOG] ", format, ap);
  pthread_mutex_unlock(&g_log_lock);
  va_end(ap);
}

void luminary_print_info(bool log, const char* format, ...) {
  va_list ap, aq;
  va_start(ap, format);
  va_copy(aq, ap);
  pthread_mutex_lock(&g_log_lock);
  console(stdout, NULL, format, ap, true);
  if (log)
    log_append("[INFO] ", format, aq);
  pthread_mutex_unlock(&g_log_lock);
  va_end(aq);
  va_end(ap);
}

void luminary_print_info_inline(bool log, const char* format, ...) {
  va_list ap, aq;
  va_start(ap, format);
  va_copy(aq, ap);
  pthread_mutex_lock(&g_log_lock);
  console(stdout, NULL, format, ap, false);
  g_inline_pending = true;
  if (log)
    log_append("[INFO] ", format, aq);
  pthread_mutex_unlock(&g_log_lock);
  va_end(aq);
  va_end(ap);
}

void luminary_print_warn(const char* format, ...) {
  va_list ap, aq;
  va_start(ap, format);
  va_copy(aq, ap);
  pthread_mutex_lock(&g_log_lock);
  console(stdout, "\033[33m", format, ap, true);
  log_append("[WARN] ", format, aq);
  pthread_mutex_unlock(&g_log_lock);
  va_end(aq);
  va_end(ap);
}

void luminary_print_error(const char* format, ...) {
  va_list ap, aq;
  va_start(ap, format);
  va_copy(aq, ap);
  pthread_mutex_lock(&g_log_lock);
  console(stderr, "\033[31m", format, ap, true);
  log_append("[ERR ] ", format, aq);
  pthread_mutex_unlock(&g_log_lock);
  va_end(aq);
  va_end(ap);
}

void luminary_write_log(void) {
  pthread_mutex_lock(&g_log_lock);
  FILE* f = fopen("luminary.log", "wb");
  if (f) {
    if (g_log_len)
      fwrite(g_log, 1, g_log_len, f);
    fclose(f);
  }
  pthread_mutex_unlock(&g_log_lock);
}

void luminary_print_crash(const char* format, ...) {
  va_list ap, aq;
  va_start(ap, format);
  va_copy(aq, ap);
  pthread_mutex_lock(&g_log_lock);
  console(stderr, "\033[35m", format, ap, true);
  log_append("[CRSH] ", format, aq);
  pthread_mutex_unlock(&g_log_lock);
  va_end(aq);
  va_end(ap);
  luminary_write_log();
  exit(EXIT_FAILURE);
}

void lum_log_shutdown(void) {
  pthread_mutex_lock(&g_log_lock);
  free(g_log);
  g_log     = NULL;
  g_log_len = g_log_cap = 0;
  pthread_mutex_unlock(&g_log_lock);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* queue: bounded ring of fixed-size elements, optional blocking pop                                            */
/* ------------------------------------------------------------------------------------------------------------ */
struct LuminaryQueue {
  pthread_mutex_t lock;
  pthread_cond_t not_empty;
  uint8_t* data;
  size_t size_of_element, capacity, head, count;
  bool is_blocking;
};

LuminaryResult _queue_create(LuminaryQueue** queue, size_t size_of_element, size_t num_elements, const char* buf_name, const char* func, uint32_t line) {
  if (!queue)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue is NULL");
  if (size_of_element == 0 || num_elements == 0)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "a queue needs a positive element size and capacity");
  LuminaryQueue* q = NULL;
  LUM_TRY(_host_malloc((void**) &q, sizeof(LuminaryQueue), buf_name, func, line));
  memset(q, 0, sizeof(*q));
  LuminaryResult r = _host_malloc((void**) &q->data, size_of_element * num_elements, buf_name, func, line);
  if (r != LUMINARY_SUCCESS) {
    _host_free((void**) &q, buf_name, func, line);
    return r | LUMINARY_ERROR_PROPAGATED;
  }
  pthread_mutex_init(&q->lock, NULL);
  pthread_cond_init(&q->not_empty, NULL);
  q->size_of_element = size_of_element;
  q->capacity        = num_elements;
  q->is_blocking     = true;
  *queue             = q;
  return LUMINARY_SUCCESS;
}

static LuminaryResult queue_push_locked(LuminaryQueue* q, void* object) {
  if (q->count == q->capacity)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "Queue ran out of memory.");
  memcpy(q->data + q->size_of_element * ((q->head + q->count) % q->capacity), object, q->size_of_element);
  q->count++;
  pthread_cond_signal(&q->not_empty);
  return LUMINARY_SUCCESS;
}

LuminaryResult queue_push(LuminaryQueue* queue, void* object) {
  if (!queue || !object)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue_push: NULL argument");
  pthread_mutex_lock(&queue->lock);
  const LuminaryResult r = queue_push_locked(queue, object);
  pthread_mutex_unlock(&queue->lock);
  return r;
}

LuminaryResult queue_push_unique(LuminaryQueue* queue, void* object, LuminaryEqOp equal_operator, bool* already_queued) {
  if (!queue || !object || !equal_operator)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue_push_unique: NULL argument");
  pthread_mutex_lock(&queue->lock);
  bool found = false;
  for (size_t k = 0; k < queue->count && !found; k++)
    found = equal_operator(queue->data + queue->size_of_element * ((queue->head + k) % queue->capacity), object);
  LuminaryResult r = LUMINARY_SUCCESS;
  if (!found)
    r = queue_push_locked(queue, object);
  pthread_mutex_unlock(&queue->lock);
  if (already_queued)
    *already_queued = found;
  return r;
}

static bool queue_pop_locked(LuminaryQueue* q, void* object) {
  if (q->count == 0)
    return false;
  memcpy(object, q->data + q->size_of_element * q->head, q->size_of_element);
  q->head = (q->head + 1) % q->capacity;
  q->count--;
  return true;
}

LuminaryResult queue_pop(LuminaryQueue* queue, void* object, bool* success) {
  if (!queue || !object || !success)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue_pop: NULL argument");
  pthread_mutex_lock(&queue->lock);
  *success = queue_pop_locked(queue, object);
  pthread_mutex_unlock(&queue->lock);
  return LUMINARY_SUCCESS;
}

/* waits for an element while the queue is in blocking mode; queue_set_is_blocking(false) releases every waiter */
LuminaryResult queue_pop_blocking(LuminaryQueue* queue, void* object, bool* success) {
  if (!queue || !object || !success)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue_pop_blocking: NULL argument");
  pthread_mutex_lock(&queue->lock);
  while (queue->count == 0 && queue->is_blocking)
    pthread_cond_wait(&queue->not_empty, &queue->lock);
  *success = queue_pop_locked(queue, object);
  pthread_mutex_unlock(&queue->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult queue_set_is_blocking(LuminaryQueue* queue, bool is_blocking) {
  if (!queue)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue is NULL");
  pthread_mutex_lock(&queue->lock);
  queue->is_blocking = is_blocking;
  pthread_cond_broadcast(&queue->not_empty);
  pthread_mutex_unlock(&queue->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult _queue_destroy(LuminaryQueue** queue, const char* buf_name, const char* func, uint32_t line) {
  if (!queue || !*queue)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "queue is NULL");
  LuminaryQueue* q = *queue;
  pthread_mutex_destroy(&q->lock);
  pthread_cond_destroy(&q->not_empty);
  LUM_TRY(_host_free((void**) &q->data, buf_name, func, line));
  LUM_TRY(_host_free((void**) queue, buf_name, func, line));
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* ring buffer: entries are carved off the head in allocation order and released from the tail in the same order */
/* ------------------------------------------------------------------------------------------------------------ */
struct LuminaryRingBuffer {
  uint8_t* memory;
  size_t size, head, tail, used;
  size_t wrap_waste; /* bytes skipped at the end of the arena by the entry that started over at the front */
};

LuminaryResult _ringbuffer_create(LuminaryRingBuffer** buffer, size_t size, const char* buf_name, const char* func, uint32_t line) {
  if (!buffer)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "buffer is NULL");
  if (size == 0)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "a ring buffer needs a positive size");
  LuminaryRingBuffer* b = NULL;
  LUM_TRY(_host_malloc((void**) &b, sizeof(LuminaryRingBuffer), buf_name, func, line));
  memset(b, 0, sizeof(*b));
  LuminaryResult r = _host_malloc((void**) &b->memory, size, buf_name, func, line);
  if (r != LUMINARY_SUCCESS) {
    _host_free((void**) &b, buf_name, func, line);
    return r | LUMINARY_ERROR_PROPAGATED;
  }
  b->size = size;
  *buffer = b;
  return LUMINARY_SUCCESS;
}

LuminaryResult ringbuffer_allocate_entry(LuminaryRingBuffer* buffer, size_t entry_size, void** entry) {
  if (!buffer || !entry)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "ringbuffer_allocate_entry: NULL argument");
  if (entry_size > buffer->size)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "Ringbuffer entry of %zu bytes exceeds the buffer (%zu bytes).", entry_size, buffer->size);
  const size_t waste = (buffer->head + entry_size > buffer->size) ? buffer->size - buffer->head : 0; /* does not fit behind the head */
  if (waste && buffer->wrap_waste)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "Ringbuffer ran out of memory.");
  if (buffer->used + waste + entry_size > buffer->size)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "Ringbuffer ran out of memory.");
  if (waste) {
    buffer->wrap_waste = waste;
    buffer->used += waste;
    buffer->head = 0;
  }
  *entry = buffer->memory + buffer->head;
  buffer->head += entry_size;
  buffer->used += entry_size;
  return LUMINARY_SUCCESS;
}

/* entries are released in allocation order; entry_size must be the size the oldest entry was allocated with */
LuminaryResult ringbuffer_release_entry(LuminaryRingBuffer* buffer, size_t entry_size) {
  if (!buffer)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "buffer is NULL");
  if (buffer->wrap_waste && buffer->tail + entry_size > buffer->size) { /* the oldest entry is the one that started over */
    buffer->used -= buffer->wrap_waste;
    buffer->wrap_waste = 0;
    buffer->tail       = 0;
  }
  if (entry_size > buffer->used)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Ringbuffer released more than was allocated.");
  buffer->tail += entry_size;
  buffer->used -= entry_size;
  if (buffer->used == 0)
    buffer->head = buffer->tail = 0;
  return LUMINARY_SUCCESS;
}

LuminaryResult _ringbuffer_destroy(LuminaryRingBuffer** buffer, const char* buf_name, const char* func, uint32_t line) {
  if (!buffer || !*buffer)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "buffer is NULL");
  LUM_TRY(_host_free((void**) &(*buffer)->memory, buf_name, func, line));
  LUM_TRY(_host_free((void**) buffer, buf_name, func, line));
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* thread status                                                                                                */
/* ------------------------------------------------------------------------------------------------------------ */
struct LuminaryThreadStatus {
  pthread_mutex_t lock;
  const char* worker_name;
  const char* string;
  double start, last_duration;
  bool running;
};

static double wall_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

LuminaryResult thread_status_create(LuminaryThreadStatus** thread_status) {
  if (!thread_status)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status is NULL");
  LuminaryThreadStatus* t = NULL;
  LUM_TRY(_host_malloc((void**) &t, sizeof(LuminaryThreadStatus), "thread_status", __func__, __LINE__));
  memset(t, 0, sizeof(*t));
  pthread_mutex_init(&t->lock, NULL);
  *thread_status = t;
  return LUMINARY_SUCCESS;
}

LuminaryResult thread_status_set_worker_name(LuminaryThreadStatus* t, const char* name) {
  if (!t)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status is NULL");
  pthread_mutex_lock(&t->lock);
  t->worker_name = name;
  pthread_mutex_unlock(&t->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult thread_status_get_worker_name(LuminaryThreadStatus* t, const char** name) {
  if (!t || !name)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status_get_worker_name: NULL argument");
  pthread_mutex_lock(&t->lock);
  *name = t->worker_name;
  pthread_mutex_unlock(&t->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult thread_status_start(LuminaryThreadStatus* t, const char* string) {
  if (!t)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status is NULL");
  pthread_mutex_lock(&t->lock);
  t->string  = string;
  t->start   = wall_seconds();
  t->running = true;
  pthread_mutex_unlock(&t->lock);
  return LUMINARY_SUCCESS;
}

/* seconds the current task has been running, or the duration of the last one once it was stopped */
LuminaryResult thread_status_get_time(LuminaryThreadStatus* t, double* time) {
  if (!t || !time)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status_get_time: NULL argument");
  pthread_mutex_lock(&t->lock);
  *time = t->running ? wall_seconds() - t->start : t->last_duration;
  pthread_mutex_unlock(&t->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult thread_status_get_string(LuminaryThreadStatus* t, const char** string) {
  if (!t || !string)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status_get_string: NULL argument");
  pthread_mutex_lock(&t->lock);
  *string = t->running ? t->string : NULL;
  pthread_mutex_unlock(&t->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult thread_status_stop(LuminaryThreadStatus* t) {
  if (!t)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status is NULL");
  pthread_mutex_lock(&t->lock);
  if (t->running)
    t->last_duration = wall_seconds() - t->start;
  t->running = false;
  pthread_mutex_unlock(&t->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult thread_status_destroy(LuminaryThreadStatus** thread_status) {
  if (!thread_status || !*thread_status)
    LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "thread_status is NULL");
  pthread_mutex_destroy(&(*thread_status)->lock);
  LUM_TRY(_host_free((void**) thread_status, "thread_status", __func__, __LINE__));
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* display names of the enumerators (reference name_strings.c)                                                  */
/* ------------------------------------------------------------------------------------------------------------ */
const char* const luminary_strings_shading_mode[LUMINARY_SHADING_MODE_COUNT] = {"None", "Albedo", "Depth", "Normal", "Identification", "Lights"};
const char* const luminary_strings_adaptive_sampling_output_mode[LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_COUNT] = {"Beauty", "Rel Variance", "Rel Error",
                                                                                                                   "Sample Distribution"};
const char* const luminary_strings_filter[LUMINARY_FILTER_COUNT]   = {"None", "Gray", "Sepia", "Gameboy", "2 Bit Gray", "CRT", "Black & White"};
const char* const luminary_strings_tonemap[LUMINARY_TONEMAP_COUNT] = {"None", "ACES", "Reinhard", "Uncharted 2", "Agx", "Agx Punchy", "Agx Custom"};
const char* const luminary_strings_aperture[LUMINARY_APERTURE_COUNT] = {"Round", "Bladed"};
const char* const luminary_strings_jerlov_water_type[LUMINARY_JERLOV_WATER_TYPE_COUNT] = {"I", "IA", "IB", "II", "III", "1C", "3C", "5C", "7C", "9C"};
const char* const luminary_strings_sky_mode[LUMINARY_SKY_MODE_COUNT] = {"Default", "HDRI", "Constant Color"};
const char* const luminary_strings_material_base_substrate[LUMINARY_MATERIAL_BASE_SUBSTRATE_COUNT] = {"Opaque", "Translucent"};
