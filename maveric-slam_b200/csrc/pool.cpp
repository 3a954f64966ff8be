// pool.cpp -- the local feature pool (reference: include/local_feature_pool.h:16-336):
// an open-addressing hash map word_id -> LocalFeature with linear probing, a
// backward-shift delete and an age-out sweep.  It is host-side bookkeeping in the
// reference and stays host-side here; the reference defines these
// functions in its header, here the header declares and this file defines, so the
// symbols link once.  Behaviour follows the reference call for call, including the
// probe order, which decides where entries end up and therefore every later lookup.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/maveric_slam_compat.h"

extern "C" {

void init_local_feature(LocalFeature* f) {  // :24-28
  f->word_id = -1;
  f->frame_ptr = 0;
  f->num_frames = 0;
}

void init_local_feature_with_id(LocalFeature* f, int word_id, int frame_num) {  // :30-35
  f->word_id = word_id;
  f->frame_ptr = 0;
  f->num_frames = 1;
  f->frames[0] = frame_num;
}

// ring buffer of the last MAX_LOCAL_FRAMES sightings (:37-47)
void update_local_feature(LocalFeature* f, int frame_num) {
  if (f->num_frames < MAX_LOCAL_FRAMES) {
    f->frames[(f->frame_ptr + f->num_frames) % MAX_LOCAL_FRAMES] = frame_num;
    f->num_frames += 1;
    return;
  }
  f->frames[f->frame_ptr] = frame_num;
  f->frame_ptr = (f->frame_ptr + 1) % MAX_LOCAL_FRAMES;
}

// drops at most one expired sighting per call; true when none is left (:49-62)
bool remove_old_frame(LocalFeature* f, int oldest_keep_frame) {
  if (f->word_id == -1) return false;
  if (f->frames[f->frame_ptr] < oldest_keep_frame) {
    f->frame_ptr = (f->frame_ptr + 1) % MAX_LOCAL_FRAMES;
    f->num_frames -= 1;
  }
  return f->num_frames == 0;
}

void init_hash_entry(HashEntry* e) {  // :70-74
  e->key = -1;
  e->is_occupied = false;
  init_local_feature(&e->value);
}

void delete_hash_entry(HashEntry* e) { init_hash_entry(e); }  // :76-80

int hash(int key, int capacity) { return key % capacity; }  // :93-95

void init_local_feature_pool(LocalFeaturePool* pool) {  // :97-103
  pool->size = 0;
  pool->capacity = LOCAL_FEATURE_POOL_CAPACITY;
  for (int i = 0; i < pool->capacity; i++) init_hash_entry(&pool->entries[i]);
}

// :108-131 -- existing key: {value, false}; free slot on the probe path: {value, true};
// full table: {NULL, false}
LocalFeaturePoolInsertResult local_feature_pool_insert(LocalFeaturePool* pool, int key, LocalFeature value) {
  LocalFeaturePoolInsertResult res = {NULL, false};
  if (pool->size >= pool->capacity) return res;
  int slot = hash(key, pool->capacity);
  for (int probes = 0; probes < pool->capacity; probes++, slot = (slot + 1) % pool->capacity) {
    HashEntry* e = &pool->entries[slot];
    if (e->key == key) {
      res.feature = &e->value;
      return res;
    }
    if (!e->is_occupied) {
      e->key = key;
      e->value = value;
      e->is_occupied = true;
      pool->size += 1;
      res.feature = &e->value;
      res.inserted = true;
      return res;
    }
  }
  return res;
}

// :137-168 -- after removing `hole`, repeatedly pull back the LAST entry of the
// following run whose home slot lies at or before the hole, until none qualifies;
// returns the slot that finally has to be cleared.
int chain_replacement(LocalFeaturePool* pool, int hole) {
  int last = hole;
  for (;;) {
    int pick = -1;
    int slot = (hole + 1) % pool->capacity;
    bool wrapped = (slot == 0);
    for (int probes = 0; probes < pool->capacity; probes++) {
      const HashEntry* e = &pool->entries[slot];
      if (!e->is_occupied) break;
      const int home = hash(e->key, pool->capacity);
      if (!wrapped) {
        if (home <= hole) pick = slot;
      } else if (home > slot && home <= hole) {
        pick = slot;
      }
      slot = (slot + 1) % pool->capacity;
      if (slot == 0) wrapped = true;
    }
    if (pick == -1) break;
    pool->entries[hole] = pool->entries[pick];
    hole = pick;
    last = pick;
  }
  return last;
}

// :170-192 -- a missing key is fatal in the reference ("Key not found", exit(0))
bool local_feature_pool_delete(LocalFeaturePool* pool, int key) {
  int slot = hash(key, pool->capacity);
  int found = -1;
  for (int probes = 0; probes < pool->capacity; probes++, slot = (slot + 1) % pool->capacity) {
    if (!pool->entries[slot].is_occupied) {
      printf("Key not found\n");
      exit(0);
    }
    if (pool->entries[slot].key == key) {
      found = slot;
      break;
    }
  }
  const int clear = chain_replacement(pool, found);
  delete_hash_entry(&pool->entries[clear]);
  pool->size -= 1;
  return true;
}

void local_feature_pool_rehash(LocalFeaturePool* pool) {  // :234-251
  HashEntry* old = (HashEntry*)malloc(sizeof(HashEntry) * (size_t)pool->capacity);
  for (int i = 0; i < pool->capacity; i++) old[i] = pool->entries[i];
  pool->size = 0;
  for (int i = 0; i < pool->capacity; i++) init_hash_entry(&pool->entries[i]);
  for (int i = 0; i < pool->capacity; i++)
    if (old[i].is_occupied) local_feature_pool_insert(pool, old[i].key, old[i].value);
  free(old);
}

float local_feature_pool_load_factor(LocalFeaturePool* pool) {  // :253-255
  return (float)pool->size / pool->capacity;
}

// :258-269 -- a deletion can pull another entry into slot i, so slot i is looked at again
void local_feature_pool_remove_old(LocalFeaturePool* pool, int current_frame_num) {
  const int keep_from = current_frame_num - MAX_LOCAL_FRAMES + 1;
  for (int i = 0; i < pool->capacity; i++) {
    HashEntry* e = &pool->entries[i];
    if (e->is_occupied && remove_old_frame(&e->value, keep_from)) {
      local_feature_pool_delete(pool, e->key);
      i--;
    }
  }
}

void local_feature_pool_valid_keys(LocalFeaturePool* pool, int* num_keys, int* keys) {  // :271-277
  for (int i = 0; i < pool->capacity; i++)
    if (pool->entries[i].is_occupied) keys[(*num_keys)++] = pool->entries[i].key;
}

// :279-336 -- the reference's self-check, same messages, exit(0) on violation
void local_feature_pool_check_invariant(LocalFeaturePool* pool, int cur_frame, bool print) {
  int counted = 0;
  for (int i = 0; i < pool->capacity; i++) {
    HashEntry* e = &pool->entries[i];
    if (!e->is_occupied) continue;
    LocalFeature* f = &e->value;
    if (print) {
      printf("Index %d, Feature %d: ", i, e->key);
      for (int j = 0, at = f->frame_ptr; j < f->num_frames; j++, at = (at + 1) % MAX_LOCAL_FRAMES)
        printf("Frame %d ", f->frames[at]);
      printf("\n");
    }
    counted++;
    if (e->key == -1) { printf("Entry key is -1\n"); exit(0); }
    if (f->word_id != e->key) {
      printf("Entry key %d does not match value word id %d\n", e->key, f->word_id);
      fflush(stdout);
      exit(0);
    }
    if (f->num_frames < 1) { printf("Feature has no frames\n"); exit(0); }
    int at = f->frame_ptr;
    if (f->frames[at] < cur_frame - MAX_LOCAL_FRAMES + 1) {
      printf("Frame %d is too old\n", f->frames[at]);
      exit(0);
    }
    for (int j = 1; j < f->num_frames; j++) {
      const int prev = at;
      at = (at + 1) % MAX_LOCAL_FRAMES;
      if (f->frames[at] <= f->frames[prev]) { printf("Frames are not in increasing order\n"); exit(0); }
    }
  }
  if (counted != pool->size) { printf("Size count is incorrect: %d %d\n", counted, pool->size); exit(0); }
  if (print) printf("Load factor: %f\n", local_feature_pool_load_factor(pool));
}

}  // extern "C"
